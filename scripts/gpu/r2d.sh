#!/bin/bash
# round 2, call D: Chronos-2 full fine-tune gradients, full fine-tune loss curve under graph replay, bench full-finetune
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chronos_gpu.py tests/test_finetune_gpu.py tests/test_parity_gpu.py tests/test_kernels_gpu.py tests/test_checkpoint_gpu.py -m gpu -q --timeout 600 -s -k "full_finetune or loss_curve or whole_stack or fallback or checkpoint or patchify" > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?"; grep -E "loss curve|worst" gpurun_out/r2d_tests.log | cut -c1-500; tail -12 gpurun_out/r2d_tests.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2d_pytest.log
for G in "" "--no-graphs"; do
  timeout 600 python bench.py --workload full-finetune --steps 10 --warmup 3 $G > gpurun_out/r2d_bench_full$G.json 2> gpurun_out/r2d_bench_full$G.err
  echo "bench full $G rc=$?"; tail -2 gpurun_out/r2d_bench_full$G.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2d_bench_full$G.json"))
    print("full $G", round(d["value"]), "series/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], "roofline", round(d["roofline"]["frac"], 3), d["clocks"])
except Exception as e:
    print("no line", e)
PY
done
python bench.py --steps 10 --warmup 3 --no-parity --no-stages --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2d_bench.json')); print('forecast', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['achieved'], d['clocks'])"
python bench.py --steps 10 --warmup 3 --no-parity --no-stages --no-cpu-baseline --no-graphs > gpurun_out/r2d_bench_eager.json 2> gpurun_out/r2d_bench_eager.err; python -c "
import json; d=json.load(open('gpurun_out/r2d_bench_eager.json')); print('forecast eager', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])"
python scripts/bench_hbm_kernels.py --iters 20 > gpurun_out/r2d_hbm_kernels.jsonl 2> gpurun_out/r2d_hbm.err; python - <<'PY'
import json
for l in open("gpurun_out/r2d_hbm_kernels.jsonl"):
    d = json.loads(l); print(d["kernel"], d["ms"], d["achieved"], d["frac"])
PY
