#!/bin/bash
# round 2, call AA (1 GPU): ncu --set full of the fine-tune step's backward kernels (attention backward, junction backward) and
# of the token-major weight-gradient GEMM
mkdir -p gpurun_out
F="python bench.py --workload full-finetune --steps 1 --warmup 2 --no-graphs --no-parity --no-stages --no-cpu-baseline"
timeout 300 $F > gpurun_out/r2aa_plain.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"timesfm_attention_bwd_mma|rmsnorm_bwd_chain" --launch-skip 220 -c 4 \
  -o gpurun_out/r2aa_bwd $F > gpurun_out/r2aa_ncu_bwd.log 2>&1
echo "bwd capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tcgen05_kernel<256, 2, true>|gemm_bf16_tcgen05_kernelILi256ELi2ELb1" --launch-skip 220 -c 4 \
  -o gpurun_out/r2aa_wgrad $F > gpurun_out/r2aa_ncu_wgrad.log 2>&1
echo "wgrad capture rc=$?"; ls -la gpurun_out/r2aa*.ncu-rep; tail -3 gpurun_out/r2aa_ncu_wgrad.log
