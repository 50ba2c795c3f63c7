#!/bin/bash
# round 2, call D: Chronos-2 full fine-tune gradients, full fine-tune loss curve under graph replay, bench full-finetune
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chronos_gpu.py tests/test_finetune_gpu.py -m gpu -q --timeout 600 -s -k "full_finetune or loss_curve" > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?"; grep -E "loss curve|worst" gpurun_out/r2d_tests.log | cut -c1-500; tail -12 gpurun_out/r2d_tests.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2d_pytest.log
for G in "" "--no-graphs"; do
  timeout 600 python bench.py --workload full-finetune --steps 10 --warmup 3 $G > gpurun_out/r2d_bench_full$G.json 2> gpurun_out/r2d_bench_full$G.err
  echo "bench full $G rc=$?"; tail -2 gpurun_out/r2d_bench_full$G.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2d_bench_full$G.json"))
    print("full $G", round(d["value"]), "series/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "roofline", round(d["roofline"]["frac"], 3), d["clocks"])
except Exception as e:
    print("no line", e)
PY
done
