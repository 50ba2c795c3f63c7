#!/bin/bash
# round 2, call J (1 GPU): decode attention with the rope table, T5 sampling, calibrated Chronos tolerances, Chronos-2 graphs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_chronos_t5_gpu.py tests/test_chronos_gpu.py -m gpu -q --timeout 600 -s > gpurun_out/r2j_tests.log 2>&1
echo "tests rc=$?"; grep -E "ratio|product" gpurun_out/r2j_tests.log | cut -c1-300 | head; tail -6 gpurun_out/r2j_tests.log | cut -c1-400
for W in longctx-timesfm chronos2; do
  timeout 900 python bench.py --workload $W --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_${W}_n1.json 2> gpurun_out/r2j_bench_${W}_n1.err
  echo "bench $W rc=$?"; tail -2 gpurun_out/r2j_bench_${W}_n1.err | cut -c1-300
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2j_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], round(d["value"]), "series/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"]),
              "roofline", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"], d["config"]["launch"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "no line", e)
PY
