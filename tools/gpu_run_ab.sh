# A/B run: new kernels' parity first (short timeouts), full gpu suite, bench in both tensor-core MLP modes, ncu of mode 2
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core" --timeout 200 > gpurun_out/pytest_tc.log 2>&1; echo "tc rc=$?" >> gpurun_out/pytest_tc.log
tail -5 gpurun_out/pytest_tc.log
timeout 900 python -m pytest tests -m gpu -q -rA --timeout 600 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|error|rc=|adam over" gpurun_out/pytest_gpu.log | tail -12
for m in 1 2; do
  NVR_MLP_MODE=$m timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m$m.json 2> gpurun_out/bench_m$m.err; echo "bench mode $m rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_m$m.json"))
print("mode $m", d["ms_per_step"], d["stage_ms_per_step"], d.get("train_step"), d.get("roofline_l1_insitu", {}).get("frac"))
PY
done
NVR_MLP_MODE=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --steps-only > gpurun_out/ncu_launch.log 2>&1
NVR_MLP_MODE=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_" -s 71 -c 17 -o gpurun_out/prof_full -f python bench.py --steps 1 --warmup 4 --steps-only > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | head -30
