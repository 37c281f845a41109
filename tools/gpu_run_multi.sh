set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 400 > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log; tail -5 gpurun_out/pytest_multi.log
for N in 1 2; do
  if [ $N = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err; fi
  echo "N=$N rc=$?"; tail -c 600 gpurun_out/scale_$N.json; tail -3 gpurun_out/scale_$N.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --scaling strong > gpurun_out/scale_2_strong.json 2> gpurun_out/scale_2_strong.err; tail -c 400 gpurun_out/scale_2_strong.json
