#!/bin/bash
# round 2, call AL (1 GPU): junction kernel with the residual row prefetched next to the GEMM output - tests, benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_chronos_gpu.py tests/test_golden_gpu.py -m gpu -q --timeout 600 -x > gpurun_out/r2al_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2al_pytest.log
python - <<'PY'
import sys, torch
sys.path.insert(0, "multimodal-timesfm_b200")
from tsfmx_b200 import ops
dev = torch.device("cuda:0")
def timeit(fn, iters=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
for rows, cols in ((65536, 1280), (99328, 768)):
    bufs = [(torch.randn(rows, cols, device=dev).bfloat16(), torch.randn(rows, cols, device=dev), torch.empty(rows, cols, device=dev),
             torch.empty(rows, cols, device=dev, dtype=torch.bfloat16)) for _ in range(3)]
    w1, w2 = torch.rand(cols, device=dev) + 0.5, torch.rand(cols, device=dev) + 0.5
    i = [0]
    def f():
        i[0] += 1
        a, x, y, yn = bufs[i[0] % 3]
        ops.norm_residual_norm(a, x, w1 if cols == 1280 else None, w2, 1e-6, y, ops.DT_BF16, yn)
    t = timeit(f)
    print(f"norm_residual_norm {rows} x {cols}: {t:.1f} us, {rows * cols * 12 / t / 1e6:.2f} TB/s", flush=True)
PY
python bench.py --no-stages > gpurun_out/r2al_bench.json 2> gpurun_out/r2al_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2al_bench.json')); print(round(d['value']), 'series/s', 'e2e', round(d['e2e']['value']), 'parity', d['parity']['ok'], 'roofline', round(d['roofline']['frac'],3), d['roofline'].get('share_of_step'), d['clocks'])"
timeout 600 python bench.py --workload chronos2 --no-cpu-baseline > gpurun_out/r2al_bench_chronos2.json 2> /dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2al_bench_chronos2.json')); print('chronos2', round(d['value']), 'series/s', d['clocks'])"
