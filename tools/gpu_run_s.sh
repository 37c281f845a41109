# sanitizer pass on small cases + the new size tests
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "c4_c5" -s > gpurun_out/pytest_c45.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_c45.log; tail -5 gpurun_out/pytest_c45.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/sanitizer_memcheck.log 2>&1; tail -6 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/sanitizer_racecheck.log 2>&1; tail -6 gpurun_out/sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_training.py -m gpu -q -x -k "network_forward_backward or deformer_backward or composite" > gpurun_out/sanitizer_train.log 2>&1; tail -6 gpurun_out/sanitizer_train.log
