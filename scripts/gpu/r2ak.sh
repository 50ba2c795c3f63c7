#!/bin/bash
# round 2, call AK (1 GPU): fused exp2 softmax numerators in the tensor-core attention kernels, additive key mask in the
# Chronos-2 encoder attention - tests, attention probe, Chronos-2 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chronos_gpu.py tests/test_finetune_gpu.py tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_golden_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/r2ak_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2ak_pytest.log
timeout 200 python scripts/attention_probe.py > gpurun_out/r2ak_attention_probe.log 2>&1; cat gpurun_out/r2ak_attention_probe.log | tail -5
for W in chronos2 longctx-chronos2; do
  timeout 600 python bench.py --workload $W --no-cpu-baseline > gpurun_out/r2ak_bench_$W.json 2> gpurun_out/r2ak_bench_$W.err
  echo "$W rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2ak_bench_$W.json')); print(round(d['value']), 'series/s', d['ms_per_step'], 'ms e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'],3), d['clocks'])"
done
