# gather in one launch: gpu suite, A/B against one launch per part (NVR_TUNE=16), full bench, ncu full capture of one step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|error|rc=|FAILED|Error" gpurun_out/pytest_gpu.log | tail -25
for t in 0 16; do
  NVR_TUNE=$t timeout 300 python bench.py --steps 20 --warmup 3 --steps-only > gpurun_out/ab_tune$t.json 2> gpurun_out/ab_tune$t.err; echo "tune $t rc=$?"; cat gpurun_out/ab_tune$t.json
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_" -s 40 -c 24 -o gpurun_out/prof_full -f python bench.py --steps 1 --warmup 3 --steps-only > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -s 46 -c 100 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --steps-only > gpurun_out/ncu_launch.log 2>&1
