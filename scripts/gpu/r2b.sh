#!/bin/bash
# round 2, call B: AR-decode tests, SiLU epilogue (shared-reciprocal sigmoid) regression + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py -m gpu -q --timeout 600 -x -s > gpurun_out/r2b_decode.log 2>&1
echo "decode rc=$?"; grep -E "rel_max|product" gpurun_out/r2b_decode.log | head -30; tail -15 gpurun_out/r2b_decode.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_decode_gpu.py > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2b_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-parity --no-stages --no-cpu-baseline"
$B > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['avg_launch_ms'], d['clocks'])"
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:gemm_bf16_tcgen05 --launch-skip 208 -c 8 --csv \
  --log-file gpurun_out/r2b_gemm_times.csv $B --steps 1 --no-graphs --lanes 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2b_gemm_times.csv | cut -d, -f 5,13- | head -20
