#!/bin/bash
# round 2, call AF (1 GPU): forward attention with packed-fp32 (FFMA2) conditioning - tests, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_golden_gpu.py tests/test_finetune_gpu.py tests/test_decode_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/r2af_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2af_pytest.log
python bench.py --no-stages > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2af_bench.json')); print(round(d['value']), 'series/s', 'e2e', round(d['e2e']['value']), 'parity', d['parity']['ok'], d['parity']['bf16'], d['parity'].get('bf16x3'), 'roofline', round(d['roofline']['frac'],3), d['roofline'].get('share_of_step'), d['clocks'])"
