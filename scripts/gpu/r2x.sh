#!/bin/bash
# round 2, call X (1 GPU): ncu launch lists of the Chronos-2 forecast, the fusion fine-tune step and the full fine-tune step
mkdir -p gpurun_out
for W in chronos2 finetune full-finetune; do
  B="python bench.py --workload $W --steps 1 --warmup 3 --no-graphs --no-parity --no-stages --no-cpu-baseline"
  timeout 600 $B > gpurun_out/r2x_plain_$W.log 2>&1
  echo "$W plain rc=$?"; tail -c 600 gpurun_out/r2x_plain_$W.log; echo
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv \
    --log-file gpurun_out/r2x_launches_$W.csv $B > gpurun_out/r2x_ncu_$W.log 2>&1
  echo "$W launch list rc=$?"; wc -l gpurun_out/r2x_launches_$W.csv
done
