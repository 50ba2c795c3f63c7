#!/bin/bash
# round 2, call G (1 GPU): the other BASELINE configs as bench workloads at N = 1 (with the CPU oracle beside them)
mkdir -p gpurun_out
for W in chronos2 longctx-chronos2 longctx-timesfm chronos-t5; do
  timeout 900 python bench.py --workload $W --steps 5 --warmup 3 > gpurun_out/r2g_bench_${W}_n1.json 2> gpurun_out/r2g_bench_${W}_n1.err
  echo "bench $W rc=$?"; tail -2 gpurun_out/r2g_bench_${W}_n1.err | cut -c1-300
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2g_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], round(d["value"]), "series/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"]),
              "roofline", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"], d["config"]["launch"], d["clocks"]["sm_mhz"], d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "no line", e)
PY
