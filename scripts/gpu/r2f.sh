#!/bin/bash
# round 2, call F (2 GPUs): data-parallel fine-tune checks and bench lines, multi-device guard, forecast at N = 2
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 300 python -m pytest tests/test_edge_cases_gpu.py -m gpu -q -k non_current_device 2>&1 | tail -3
for mode in multimodal baseline; do timeout 300 $T 29511 scripts/ddp_finetune_check.py $mode 2>&1 | grep -E "mode=|ok|Error|error" | tail -3; done
for W in finetune full-finetune; do
  timeout 600 $T 29512 bench.py --gpus 2 --workload $W --steps 10 --warmup 3 > gpurun_out/r2f_bench_${W}_n2.json 2> gpurun_out/r2f_bench_${W}_n2.err
  echo "bench $W n2 rc=$?"; tail -2 gpurun_out/r2f_bench_${W}_n2.err | cut -c1-300
done
timeout 400 $T 29513 bench.py --gpus 2 --workload full-finetune --steps 10 --warmup 3 --graph-collectives > gpurun_out/r2f_bench_full-finetune_graphcoll_n2.json 2> gpurun_out/r2f_bench_full-finetune_graphcoll_n2.err
echo "bench full graph-collectives n2 rc=$?"; tail -3 gpurun_out/r2f_bench_full-finetune_graphcoll_n2.err | cut -c1-400
NCCL_DEBUG=INFO timeout 600 $T 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
echo "bench forecast n2 rc=$?"; grep -E "NVLS|Connected all|via P2P" gpurun_out/r2f_bench_n2.err | head -3
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2f_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], round(d["value"]), "series/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"]),
              d["config"].get("collective"), d["config"].get("launch"), d["clocks"])
    except Exception as e:
        print(f, "no line", e)
PY
