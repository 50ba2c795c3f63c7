#!/bin/bash
# round 2, call O (1 GPU): full GPU suite on the final code, the driver's bench command, ncu --set full of the kernels changed
# this round (patchify with the fast path, ff0 with the shared-reciprocal SiLU, decode attention v3)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2o_pytest.log
python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2o_bench.err | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2o_bench_reference.json 2> /dev/null; cut -c1-200 gpurun_out/r2o_bench_reference.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2o_bench.json").read())
for k in ("value", "ms_per_step", "e2e", "e2e_forecast_readback", "value_bf16x3", "parity", "clocks", "cpu_baseline", "roofline_stages_clocks"):
    print(k, json.dumps(d.get(k))[:500])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "share_of_step", "avg_launch_ms", "traffic")})
for s in d.get("roofline_stages", []):
    print(s["kernel"], s["achieved"], s["frac"])
PY
H="python scripts/bench_hbm_kernels.py --iters 1 --only bf16-out --contexts 512"
$H > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:timesfm_patchify_norm_warp --launch-skip 3 -c 1 -o gpurun_out/r2o_patchify $H > gpurun_out/r2o_ncu_patchify.log 2>&1
echo "patchify capture rc=$?"
G="python bench.py --steps 1 --warmup 3 --no-graphs --no-parity --no-stages --no-cpu-baseline --lanes 1"
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 --launch-skip 208 -c 4 -o gpurun_out/r2o_gemm $G > gpurun_out/r2o_ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
D="python bench.py --workload longctx-timesfm --steps 1 --warmup 3 --no-cpu-baseline --batch 1024"
ncu --set full --clock-control none --import-source on -k regex:attention_decode --launch-skip 210 -c 1 -o gpurun_out/r2o_decode_attention $D > gpurun_out/r2o_ncu_decode.log 2>&1
echo "decode capture rc=$?"; ls -la gpurun_out/*.ncu-rep | tail -4
