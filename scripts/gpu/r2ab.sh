#!/bin/bash
# round 2, call AB (2 GPUs): data-parallel checks and fine-tune benches after the token-major weight gradients
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 240 $T 29531 scripts/ddp_finetune_check.py > gpurun_out/r2ab_ddp_fusion.log 2>&1; echo "ddp fusion rc=$?"; tail -3 gpurun_out/r2ab_ddp_fusion.log | cut -c1-300
timeout 240 $T 29532 scripts/ddp_finetune_check.py baseline > gpurun_out/r2ab_ddp_baseline.log 2>&1; echo "ddp baseline rc=$?"; tail -3 gpurun_out/r2ab_ddp_baseline.log | cut -c1-300
for W in finetune full-finetune; do
  timeout 300 $T 29533 bench.py --gpus 2 --workload $W > gpurun_out/r2ab_bench_${W}_n2.json 2> gpurun_out/r2ab_bench_${W}_n2.err
  echo "$W rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2ab_bench_${W}_n2.json')); print(round(d['value']), 'series/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value']), d['config'].get('collective'), d['clocks'])"
done
