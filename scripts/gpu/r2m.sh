#!/bin/bash
# round 2, call M (1 GPU): rewritten decode attention (tests, cfg-5 bench, launch list of one cfg-5 step)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_checkpoint_gpu.py -m gpu -q --timeout 600 > gpurun_out/r2m_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r2m_tests.log | cut -c1-400
B="python bench.py --workload longctx-timesfm --steps 4 --warmup 3 --no-cpu-baseline"
timeout 900 $B > gpurun_out/r2m_bench_longctx-timesfm_n1.json 2> gpurun_out/r2m_bench_longctx-timesfm_n1.err
echo "bench rc=$?"; tail -2 gpurun_out/r2m_bench_longctx-timesfm_n1.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench_longctx-timesfm_n1.json')); print(round(d['value']), 'series/s', round(d['ms_per_step'],1), 'ms', 'roofline', round(d['roofline']['frac'],3), d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4500 -c 1500 --csv --log-file gpurun_out/r2m_launches_longctx.csv $B --steps 1 > gpurun_out/r2m_ncu.log 2>&1
echo "launch list rc=$?"; python scripts/summarize_launches.py gpurun_out/r2m_launches_longctx.csv | head -14
