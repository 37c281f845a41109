set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "tensor_core" -s > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc.log
tail -30 gpurun_out/pytest_tc.log
