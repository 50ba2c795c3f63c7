#!/bin/bash
# round 2, call I (8 GPUs): BASELINE configs at their stated scale - forecast (configs[1] x 8), fine-tune steps with the NCCL
# all-reduce (configs[3]), Chronos-2 / Chronos-T5 (configs[2]), ctx 2048 / h 256 (configs[4])
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
port=29600
run() {  # run <name> <timeout> <bench args...>
  name=$1; limit=$2; shift 2; port=$((port + 1))
  timeout $limit $T $port bench.py --gpus 8 "$@" > gpurun_out/r2i_${name}_n8.json 2> gpurun_out/r2i_${name}_n8.err
  echo "$name rc=$? $(date +%T)"
}
date +%T
run forecast 300 --steps 10 --warmup 3
run finetune 240 --workload finetune --steps 10 --warmup 3
run full-finetune 240 --workload full-finetune --steps 10 --warmup 3
run chronos2 200 --workload chronos2 --steps 5 --warmup 3
run longctx-timesfm 240 --workload longctx-timesfm --steps 4 --warmup 3
run longctx-chronos2 200 --workload longctx-chronos2 --steps 5 --warmup 3
run chronos-t5 300 --workload chronos-t5 --steps 3 --warmup 3
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2i_*_n8.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], round(d["value"]), "series/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"]),
              "roofline", round(d["roofline"]["frac"], 3), d["config"].get("collective"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(f, "no line", e)
PY
grep -h -E "Error|error" gpurun_out/r2i_*_n8.err | sort | uniq -c | head -5
