set -x
mkdir -p gpurun_out
timeout 600 python tools/train_breakdown.py > gpurun_out/train_breakdown.log 2>&1; echo "rc=$?"; head -12 gpurun_out/train_breakdown.log; grep -A32 "Self CPU %" gpurun_out/train_breakdown.log | cut -c1-70,130-260 | head -40
