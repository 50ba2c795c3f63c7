#!/bin/bash
# round 2, call AE (1 GPU): driver-style sequence on the current tree - smoke, full GPU suite, bench
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2ae_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2ae_pytest.log
python bench.py > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2ae_bench.json')); print(round(d['value']), 'series/s', 'e2e', round(d['e2e']['value']), 'parity ok', d['parity']['ok'], d['parity']['bf16']['ratio_to_bf16_oracle'], 'roofline', round(d['roofline']['frac'],3), d['clocks'])
for s in d.get('roofline_stages', []): print(s['kernel'], s['achieved'], s['frac'])"
