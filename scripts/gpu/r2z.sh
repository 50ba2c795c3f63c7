#!/bin/bash
# round 2, call Z (1 GPU): token-major weight gradients + scale gradients fused into the junction backward - tests, fine-tune benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_finetune_gpu.py tests/test_chronos_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/r2z_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2z_pytest.log
for W in finetune full-finetune; do
  timeout 600 python bench.py --workload $W --no-cpu-baseline > gpurun_out/r2z_bench_$W.json 2> gpurun_out/r2z_bench_$W.err
  echo "$W rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2z_bench_$W.json')); print(round(d['value']), 'series/s', d['ms_per_step'], 'ms e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'],3), d['clocks'])"
done
