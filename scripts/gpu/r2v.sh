#!/bin/bash
# round 2, call V (1 GPU): decode attention v4 (five lanes per key row) - tests, cfg-5 bench, launch timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_checkpoint_gpu.py -m gpu -q --timeout 600 > gpurun_out/r2v_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r2v_tests.log | cut -c1-400
B="python bench.py --workload longctx-timesfm --steps 6 --warmup 6 --no-cpu-baseline"
timeout 900 $B > gpurun_out/r2v_bench_longctx-timesfm_n1.json 2> gpurun_out/r2v_bench_longctx-timesfm_n1.err
echo "bench rc=$?"; tail -2 gpurun_out/r2v_bench_longctx-timesfm_n1.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2v_bench_longctx-timesfm_n1.json')); print(round(d['value']), 'series/s', round(d['ms_per_step'],1), 'ms', 'e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'],3), d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attention_decode --launch-skip 300 -c 40 --csv --log-file gpurun_out/r2v_decode_times.csv $B --steps 1 --warmup 3 > gpurun_out/r2v_ncu.log 2>&1
grep -v "^==" gpurun_out/r2v_decode_times.csv | tail -3 | cut -d, -f 5,12-
