#!/bin/bash
# round 2, call Y (1 GPU): token-major weight-gradient GEMM - descriptor variants, accuracy, timing against the transposing path
mkdir -p gpurun_out
timeout 300 python scripts/wgrad_probe.py --variants > gpurun_out/r2y_wgrad_probe.log 2>&1
echo "probe rc=$?"; cat gpurun_out/r2y_wgrad_probe.log | tail -40
