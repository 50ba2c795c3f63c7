#!/bin/bash
# round 2, call AM (1 GPU): driver-style sequence on the final tree (smoke, full GPU suite, bench, reference arm), then the
# ncu launch list of the forecast step as it stands
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2am_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2am_pytest.log
python bench.py > gpurun_out/r2am_bench.json 2> gpurun_out/r2am_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2am_bench.json')); print(round(d['value']), 'series/s', 'e2e', round(d['e2e']['value']), 'parity ok', d['parity']['ok'], d['parity']['bf16']['ratio_to_bf16_oracle'], d['parity']['bf16x3']['rel_max'], 'roofline', round(d['roofline']['frac'],3), d['roofline'].get('share_of_step'), d['clocks'])
for s in d.get('roofline_stages', []): print(s['kernel'], s['achieved'], s['frac'])"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2am_bench_reference.json 2> /dev/null; cut -c1-260 gpurun_out/r2am_bench_reference.json
B="python bench.py --steps 1 --warmup 3 --no-graphs --no-parity --no-stages --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r2am_launches.csv $B > gpurun_out/r2am_ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2am_launches.csv
