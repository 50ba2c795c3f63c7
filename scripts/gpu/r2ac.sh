#!/bin/bash
# round 2, call AC (1 GPU): TimesFM patchify at ctx 1024 / 2048 in bf16 - launch-shape variants, ncu of the ctx 2048 case
mkdir -p gpurun_out
H="python scripts/bench_hbm_kernels.py --iters 20 --only timesfm"
$H --contexts 1024,2048,4096 > gpurun_out/r2ac_default.jsonl 2>&1
$H --contexts 1024 --tune 0=1 > gpurun_out/r2ac_ctx1024_g1.jsonl 2>&1
$H --contexts 2048 --tune 1=520 > gpurun_out/r2ac_ctx2048_2stages.jsonl 2>&1
$H --contexts 2048 --tune 1=4 > gpurun_out/r2ac_ctx2048_4warps.jsonl 2>&1
for f in default ctx1024_g1 ctx2048_2stages ctx2048_4warps; do echo "== $f"; python -c "
import json,sys
for l in open('gpurun_out/r2ac_$f.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['kernel'], d['ms'], d['frac'])"; done
N="python scripts/bench_hbm_kernels.py --iters 1 --only bf16-out --contexts 2048"
ncu --set full --clock-control none --import-source on -k regex:timesfm_patchify_norm_warp --launch-skip 3 -c 1 -o gpurun_out/r2ac_patchify2048 $N > gpurun_out/r2ac_ncu.log 2>&1
echo "capture rc=$?"; ls -la gpurun_out/r2ac_patchify2048.ncu-rep
