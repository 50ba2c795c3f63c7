#!/bin/bash
# round 2, call R (1 GPU): per-kernel budget of a Chronos-T5 greedy decode step at 2048 series
mkdir -p gpurun_out
python scripts/t5_decode_probe.py 2048 8
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 1200 --csv --log-file gpurun_out/r2r_launches_t5_decode.csv python scripts/t5_decode_probe.py 2048 8 > gpurun_out/r2r_ncu.log 2>&1
echo "launch list rc=$?"; python scripts/summarize_launches.py gpurun_out/r2r_launches_t5_decode.csv > gpurun_out/r2r_launches_t5_decode.md; cat gpurun_out/r2r_launches_t5_decode.md
python - <<'PY'
import csv, collections
rows = [r for r in csv.DictReader(l for l in open("gpurun_out/r2r_launches_t5_decode.csv") if not l.startswith("=="))]
# durations of the t5_attention launches: cross-attention (513 keys) and self-attention alternate
att = [float(r["Metric Value"]) / 1e3 for r in rows if "t5_attention_kernel" in r["Kernel Name"]]
print("t5_attention launches:", len(att), "first 12 (us):", [round(a) for a in att[:12]])
gem = [float(r["Metric Value"]) / 1e3 for r in rows if "gemm_bf16" in r["Kernel Name"]]
print("gemm launches:", len(gem), "first 16 (us):", [round(a, 1) for a in gem[:16]])
PY
