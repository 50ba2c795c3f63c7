#!/bin/bash
# round 2, call C: loss-curve tests (eager + graph replay), fine-tune workloads of bench.py on one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_finetune_gpu.py -m gpu -q --timeout 600 -x -s -k "loss_curve or oracle" > gpurun_out/r2c_finetune.log 2>&1
echo "finetune tests rc=$?"; grep -E "loss curve" gpurun_out/r2c_finetune.log | cut -c1-400; tail -12 gpurun_out/r2c_finetune.log
for W in finetune full-finetune; do
  for G in "" "--no-graphs"; do
    timeout 600 python bench.py --workload $W --steps 10 --warmup 3 $G > gpurun_out/r2c_bench_${W}${G}.json 2> gpurun_out/r2c_bench_${W}${G}.err
    echo "bench $W $G rc=$?"; tail -2 gpurun_out/r2c_bench_${W}${G}.err | cut -c1-300
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2c_bench_${W}${G}.json"))
    print("$W $G", round(d["value"]), "series/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "roofline", round(d["roofline"]["frac"], 3), d["clocks"], d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("no line", e)
PY
  done
done
