#!/bin/bash
# One parametrised GPU-box script (replaces the 17 single-purpose tools/gpu_run_*.sh of round 1).
#   gpurun --timeout 1500 -- 'bash tools/gpu_run.sh tests smoke bench ref ncu_list ncu_full train'
# stages:  tests | smoke | bench | ref | refgpu | ncu_list | ncu_full | ncu_src:<kernel-regex> | train | san | ab:<tune>[,<tune>...]
#          | multi:<N>  (pytest tests/test_gpu_multi.py + torchrun bench at N ranks)   | cfg:<c4|c5>
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for stage in "$@"; do
case "$stage" in
  tests)
    timeout 1500 python -m pytest tests -m gpu -q -rA --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
    grep -E "passed|failed|error|rc=|FAILED|Error" gpurun_out/pytest_gpu.log | tail -25 ;;
  smoke)
    timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log ;;
  bench)
    timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err ;;
  ref)
    timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_ref.json ;;
  refgpu)
    timeout 900 python bench.py --impl reference-gpu --steps 3 --warmup 1 > gpurun_out/bench_refgpu.json 2> gpurun_out/bench_refgpu.err; echo "refgpu rc=$?"; cut -c1-600 gpurun_out/bench_refgpu.json ;;
  ncu_list)
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"^k_" --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 2 --warmup 3 --steps-only > gpurun_out/ncu_launch.log 2>&1
    python tools/launch_summary.py gpurun_out/launches.csv gpurun_out/launch_list_summary.csv; cat gpurun_out/launch_list_summary.csv ;;
  ncu_full)
    timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"^k_" -o gpurun_out/prof_full -f \
        python bench.py --steps 1 --warmup 3 --steps-only > gpurun_out/ncu_full.log 2>&1
    python tools/ncu_summary.py gpurun_out/prof_full.ncu-rep gpurun_out/ncu_full_summary.csv 1; tail -2 gpurun_out/ncu_full.log ;;
  ncu_src:*)
    k="${stage#ncu_src:}"
    timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$k" -c 6 -o "gpurun_out/prof_$k" -f \
        python bench.py --steps 1 --warmup 3 --steps-only > "gpurun_out/ncu_$k.log" 2>&1; tail -2 "gpurun_out/ncu_$k.log" ;;
  san)
    for tool in memcheck synccheck; do
      timeout 420 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/sanitizer_$tool.txt 2>&1; echo "sanitizer $tool rc=$?"; tail -2 gpurun_out/sanitizer_$tool.txt
    done ;;
  train)
    timeout 300 python tools/train_breakdown.py > gpurun_out/train_breakdown.log 2>&1; head -9 gpurun_out/train_breakdown.log ;;
  ab:*)
    for t in $(echo "${stage#ab:}" | tr ',' ' '); do
      NVR_TUNE=$t timeout 600 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/ab_tune$t.json 2> gpurun_out/ab_tune$t.err
      echo "tune $t rc=$?"; python -c "import json,sys; d=json.load(open('gpurun_out/ab_tune$t.json')); print('tune $t', d['ms_per_step'], d['stage_ms_per_step'], d.get('embed_part_ms'))"
    done ;;
  mode:*)
    m="${stage#mode:}"
    NVR_MLP_MODE=$m timeout 600 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/ab_mode$m.json 2> gpurun_out/ab_mode$m.err
    echo "mlp_mode $m rc=$?"; python -c "import json,sys; d=json.load(open('gpurun_out/ab_mode$m.json')); print('mlp_mode $m', d['ms_per_step'], d['stage_ms_per_step'], d.get('mlp_part_ms'))" ;;
  cfg:*)
    c="${stage#cfg:}"
    timeout 900 python bench.py --config $c --steps 5 --warmup 3 --no-extras > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "cfg $c rc=$?"; cut -c1-800 gpurun_out/bench_$c.json ;;
  multi:*)
    n="${stage#multi:}"
    timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rA --timeout 600 > gpurun_out/pytest_multi.log 2>&1; echo "multi pytest rc=$?"; tail -5 gpurun_out/pytest_multi.log
    for mode in peer nccl; do
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 --assemble $mode \
          > gpurun_out/scale_${n}gpu_$mode.json 2> gpurun_out/scale_${n}gpu_$mode.err; echo "scale $n $mode rc=$?"
      python -c "import json; d=json.load(open('gpurun_out/scale_${n}gpu_$mode.json')); print('N=$n $mode', d['value']/1e9, 'G/s', d['ms_per_step'], 'ms; e2e', d['e2e']['value']/1e9, d['e2e']['ms_per_step'], 'c4', d.get('c4',{}).get('value'), d.get('c4',{}).get('ms_per_frame'), 'c5', d.get('c5',{}).get('value'), d.get('c5',{}).get('ms_per_frame'))"
      tail -3 gpurun_out/scale_${n}gpu_$mode.err
    done ;;
  *) echo "unknown stage $stage" ;;
esac
done
ls -la gpurun_out | head -50
