#!/bin/bash
# round 2, call AH (1 GPU): event timing of the attention kernels
mkdir -p gpurun_out
timeout 200 python scripts/attention_probe.py > gpurun_out/r2ah_attention_probe.log 2>&1; echo "rc=$?"; cat gpurun_out/r2ah_attention_probe.log | tail -8
