#!/bin/bash
# round 2, call AJ (1 GPU): ncu --set full with source of the Chronos-2 encoder attention (T = 97)
mkdir -p gpurun_out
C="python bench.py --workload chronos2 --steps 1 --warmup 1 --no-graphs --no-parity --no-stages --no-cpu-baseline"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encoder_attention_mma --launch-skip 14 -c 1 \
  -o gpurun_out/r2aj_enc_attn $C > gpurun_out/r2aj_ncu.log 2>&1
echo "capture rc=$?"; ls -la gpurun_out/r2aj_enc_attn.ncu-rep
