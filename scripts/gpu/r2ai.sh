#!/bin/bash
# round 2, call AI (1 GPU): attention backward with packed-fp32 conditioning - gradient tests, timing probe, fine-tune bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_finetune_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 600 -x -k "finetune or attention or fine_tune or loss or gradient" > gpurun_out/r2ai_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2ai_pytest.log
timeout 200 python scripts/attention_probe.py > gpurun_out/r2ai_attention_probe.log 2>&1; cat gpurun_out/r2ai_attention_probe.log | tail -5
for W in finetune; do
  timeout 600 python bench.py --workload $W --no-cpu-baseline > gpurun_out/r2ai_bench_$W.json 2> gpurun_out/r2ai_bench_$W.err
  echo "$W rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2ai_bench_$W.json')); print(round(d['value']), 'series/s', d['ms_per_step'], 'ms e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'],3), d['clocks'])"
done
