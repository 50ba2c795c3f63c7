#!/bin/bash
# round 2, call A: full GPU test suite, the driver's bench line, ncu launch list, ncu --set full of the decoder GEMMs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "rel_max|ratio" gpurun_out/r2a_pytest.log | head -40; tail -8 gpurun_out/r2a_pytest.log
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2a_bench.json").read())
for k in ("value", "ms_per_step", "e2e", "e2e_forecast_readback", "value_bf16x3", "parity", "clocks", "cpu_baseline"):
    print(k, json.dumps(d.get(k))[:700])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "share_of_step", "avg_launch_ms", "traffic")})
for s in d.get("roofline_stages", []):
    print(s["kernel"], s["achieved"], s["frac"])
PY
B="python bench.py --steps 1 --warmup 3 --no-graphs --no-parity --no-stages --no-cpu-baseline"
$B > gpurun_out/r2a_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv \
  --log-file gpurun_out/r2a_launches.csv $B > gpurun_out/r2a_ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2a_launches.csv
G="$B --lanes 1"
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 --launch-skip 208 -c 8 \
  -o gpurun_out/r2a_gemm $G > gpurun_out/r2a_ncu_gemm.log 2>&1
echo "gemm capture rc=$?"; ls -la gpurun_out/r2a_gemm.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"norm_residual_norm|timesfm_attention_mma" --launch-skip 150 -c 4 \
  -o gpurun_out/r2a_norm_attn $G > gpurun_out/r2a_ncu_norm.log 2>&1
echo "norm/attn capture rc=$?"
