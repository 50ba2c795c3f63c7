#!/bin/bash
# round 2, call U (1 GPU): cross-attention decode kernel with one block per series (all heads)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chronos_t5_gpu.py -m gpu -q --timeout 600 > gpurun_out/r2u_t5.log 2>&1
echo "t5 tests rc=$?"; tail -4 gpurun_out/r2u_t5.log | cut -c1-400
python scripts/t5_decode_probe.py 2048 8
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 500 -c 300 --csv --log-file gpurun_out/r2u_launches.csv python scripts/t5_decode_probe.py 2048 8 > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/r2u_launches.csv | head -8
timeout 900 python bench.py --workload chronos-t5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_chronos-t5_n1.json 2> gpurun_out/r2u_bench_chronos-t5_n1.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2u_bench_chronos-t5_n1.json')); print(round(d['value']), 'series/s', round(d['ms_per_step'],1), 'ms', 'roofline', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"
