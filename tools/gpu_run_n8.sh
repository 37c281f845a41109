set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err; echo "N=$N rc=$?"; tail -c 300 gpurun_out/scale_$N.json; tail -2 gpurun_out/scale_$N.err
done
