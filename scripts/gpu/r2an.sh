#!/bin/bash
# round 2, call AN (1 GPU, the last 2 GPU-minutes of the round): ncu --set full of one decoder layer's seven launches as the
# FINAL tree issues them (qkv GEMM, attention, out GEMM, junction, ff0 GEMM, ff1 GEMM, junction; 4096 series, M = 65 536)
mkdir -p gpurun_out
P="python scripts/layer_capture.py"
$P > gpurun_out/r2an_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/r2an_plain.log
timeout 90 ncu --set full --clock-control none --import-source on \
  -k regex:'gemm_bf16_tcgen05|norm_residual_norm|timesfm_attention_mma' --launch-skip 22 -c 7 \
  -o gpurun_out/r2an_layer $P > gpurun_out/r2an_ncu.log 2>&1
echo "capture rc=$?"; tail -3 gpurun_out/r2an_ncu.log; ls -la gpurun_out/r2an_layer.ncu-rep
