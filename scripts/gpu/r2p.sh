#!/bin/bash
# round 2, call P (2 GPUs): full fine-tune with the overlapped all-reduces captured into the step's CUDA graph - does the
# process now exit cleanly (graphs dropped before the communicator)?
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
date +%T
timeout 240 $T 29521 bench.py --gpus 2 --workload full-finetune --steps 10 --warmup 3 --graph-collectives > gpurun_out/r2p_full_graphcoll_n2.json 2> gpurun_out/r2p_full_graphcoll_n2.err
echo "graph-collectives rc=$? $(date +%T)"; tail -3 gpurun_out/r2p_full_graphcoll_n2.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2p_full_graphcoll_n2.json')); print(round(d['value']), round(d['ms_per_step'],2), d['config']['launch'], d['config']['collective'])"
