# quick: gpu suite, steps-only A/B of the occupancy variants, full bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|error|rc=|FAILED|Error" gpurun_out/pytest_gpu.log | tail -25
for t in 0 1 2 3; do
  NVR_TUNE=$t timeout 300 python bench.py --steps 20 --warmup 3 --steps-only > gpurun_out/ab_tune$t.json 2> gpurun_out/ab_tune$t.err; echo "tune $t rc=$?"; cat gpurun_out/ab_tune$t.json
done
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
