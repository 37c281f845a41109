# round-end style run: gpu tests, smoke, bench (with cpu baseline + train step), reference arm, ncu launch list + full capture, train breakdown
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -rA --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|error|rc=|FAILED|Error" gpurun_out/pytest_gpu.log | tail -25
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 300 python tools/train_breakdown.py > gpurun_out/train_breakdown.log 2>&1; head -9 gpurun_out/train_breakdown.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -s 58 -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --steps-only > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_" -s 58 -c 17 -o gpurun_out/prof_full -f python bench.py --steps 1 --warmup 3 --steps-only > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -40
