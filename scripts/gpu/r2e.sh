#!/bin/bash
# round 2, call E: re-run what failed in call D + evaluator.predict test, then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chronos_gpu.py tests/test_finetune_gpu.py tests/test_parity_gpu.py tests/test_kernels_gpu.py tests/test_evaluator_gpu.py -m gpu -q --timeout 600 -s -k "chronos2_match or colsum or whole_stack or patchify or predict" > gpurun_out/r2e_tests.log 2>&1
echo "tests rc=$?"; grep -E "worst|^E  " gpurun_out/r2e_tests.log | cut -c1-600 | head -20; tail -6 gpurun_out/r2e_tests.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2e_pytest.log
