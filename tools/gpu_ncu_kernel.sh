# usage: bash tools/gpu_ncu_kernel.sh <kernel-regex> <skip> <count> <outname>
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s $2 -c $3 -o gpurun_out/$4 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$4.log 2>&1
tail -5 gpurun_out/ncu_$4.log
