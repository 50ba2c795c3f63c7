#!/bin/bash
# round 2, call AD (1 GPU): patchify with the shuffle scan of the running statistics (every context)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_decode_gpu.py tests/test_edge_cases_gpu.py -m gpu -q --timeout 600 -x -k "patchify or parity or decode or edge or fallback" > gpurun_out/r2ad_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2ad_pytest.log
python scripts/bench_hbm_kernels.py --iters 20 --only timesfm --contexts 512,1024,2048,4096 > gpurun_out/r2ad_hbm.jsonl 2>&1
python -c "
import json
for l in open('gpurun_out/r2ad_hbm.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['kernel'], d['ms'], d['frac'])"
