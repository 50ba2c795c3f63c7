#!/bin/bash
# round 2, call Q (8 GPUs): the two workloads whose code changed since call I - full fine-tune (all-reduces captured into the
# step's graph) and ctx 2048 / h 256 through the rewritten decode attention
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
date +%T
timeout 240 $T 29701 bench.py --gpus 8 --workload full-finetune --steps 10 --warmup 3 > gpurun_out/r2q_full-finetune_n8.json 2> gpurun_out/r2q_full-finetune_n8.err
echo "full-finetune rc=$? $(date +%T)"
timeout 240 $T 29702 bench.py --gpus 8 --workload longctx-timesfm --steps 6 --warmup 6 > gpurun_out/r2q_longctx-timesfm_n8.json 2> gpurun_out/r2q_longctx-timesfm_n8.err
echo "longctx-timesfm rc=$? $(date +%T)"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2q_*_n8.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], round(d["value"]), "series/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"]),
              "roofline", round(d["roofline"]["frac"], 3), d["config"].get("launch"), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "no line", e)
PY
tail -2 gpurun_out/r2q_full-finetune_n8.err | cut -c1-300
