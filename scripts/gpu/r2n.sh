#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_checkpoint_gpu.py tests/test_evaluator_gpu.py tests/test_chronos_t5_gpu.py tests/test_finetune_gpu.py -m gpu -q --timeout 600 > gpurun_out/r2n_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r2n_tests.log | cut -c1-400
B="python bench.py --workload longctx-timesfm --steps 6 --warmup 6 --no-cpu-baseline"
timeout 900 $B > gpurun_out/r2n_bench_longctx-timesfm_n1.json 2> gpurun_out/r2n_bench_longctx-timesfm_n1.err
echo "bench rc=$?"; tail -2 gpurun_out/r2n_bench_longctx-timesfm_n1.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_longctx-timesfm_n1.json')); print(round(d['value']), 'series/s', round(d['ms_per_step'],1), 'ms', 'roofline', round(d['roofline']['frac'],3), d['clocks'])"
