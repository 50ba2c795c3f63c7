#!/bin/bash
# round 2, call H (1 GPU): T5 sampling kernel tests, cfg-5 TimesFM AR decode workload, T5 workload after the kernel change
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chronos_t5_gpu.py -m gpu -q --timeout 600 > gpurun_out/r2h_t5.log 2>&1
echo "t5 tests rc=$?"; tail -8 gpurun_out/r2h_t5.log | cut -c1-400
for W in longctx-timesfm; do
  timeout 900 python bench.py --workload $W --steps 4 --warmup 3 > gpurun_out/r2g_bench_${W}_n1.json 2> gpurun_out/r2g_bench_${W}_n1.err
  echo "bench $W rc=$?"; tail -2 gpurun_out/r2g_bench_${W}_n1.err | cut -c1-300
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2g_bench_longctx-timesfm*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], round(d["value"]), "series/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"]),
              "roofline", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"], d["config"]["launch"], d["clocks"]["sm_mhz"], d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "no line", e)
PY
