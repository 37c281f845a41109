# 2-GPU: multi-GPU tests + weak / strong scaling bench lines
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rA --timeout 500 > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/scale_2gpu.json 2> gpurun_out/scale_2gpu.err; echo "rc=$?"; cat gpurun_out/scale_2gpu.json; tail -2 gpurun_out/scale_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong > gpurun_out/scale_2gpu_strong.json 2> gpurun_out/scale_2gpu_strong.err; echo "rc=$?"; cat gpurun_out/scale_2gpu_strong.json
