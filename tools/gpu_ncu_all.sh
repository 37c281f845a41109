# ncu --set full over one render pass worth of our kernels (after the warm-up step), + launch list of a short bench
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_" -s ${1:-60} -c ${2:-15} -o gpurun_out/${3:-prof_all} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${3:-prof_all}.log 2>&1
tail -3 gpurun_out/ncu_${3:-prof_all}.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -s ${1:-60} -c 60 --csv --log-file gpurun_out/launches_${3:-prof_all}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
