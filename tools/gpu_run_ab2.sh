# A/B run 2: gpu suite, then the occupancy / early-out variants (NVR_TUNE bits), then a full bench of the default
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA --timeout 600 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|error|rc=|adam over|FAILED" gpurun_out/pytest_gpu.log | tail -12
for t in 0 4 1 2 3; do
  NVR_TUNE=$t timeout 300 python bench.py --steps 20 --warmup 3 --steps-only > gpurun_out/ab_tune$t.json 2> gpurun_out/ab_tune$t.err; echo "tune $t rc=$?"; cat gpurun_out/ab_tune$t.json
done
NVR_TUNE=3 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 250 > gpurun_out/pytest_tune3.log 2>&1; echo "tune3 parity rc=$?"; tail -3 gpurun_out/pytest_tune3.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
