"""BASELINE.json configs 3-5 next to the headline bench (bench.py = configs[1]); one JSON line per config.

    python scripts/bench_configs.py [--only chronos2|chronos_t5|finetune|longctx] [--steps 5] [--warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_configs.py ...

  chronos2  cfg-3: Chronos-2 (12 blocks x 768, the adapter the reference wraps) + 1-layer fusion, ctx 512 / h 128,
            2048 series per GPU, series-sharded, no collective
  chronos_t5  cfg-3 as BASELINE.json words it: Chronos-T5-base tokenise + encoder + greedy decoding of 64 tokens,
            per-token text fusion, 2048 series per GPU
  finetune  cfg-4: TimesFM "500M shape" (50 layers) multimodal fine-tune step = forward + activation-gradient pass +
            fusion weight gradient + NCCL all-reduce of the fusion gradients + clip + AdamW, 1024 series per GPU
  longctx   cfg-5: ctx 2048; TimesFM 50 layers at h 128 (the adapter API raises above 128) and Chronos-2 at h 256,
            2048 series per GPU (16k over 8 GPUs), distinct text rows per series

Every rank owns its own shard (weak scaling); time = CUDA events, max over ranks; inputs resident in HBM.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for _p in (ROOT / "multimodal-timesfm_b200", ROOT):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import timesfm_oracle as O  # noqa: E402  (synthetic Time-MMD-shaped batches only)
from tsfmx_b200 import _lib  # noqa: E402
from tsfmx_b200 import distributed as tdist  # noqa: E402
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.tsfm import chronos as C2  # noqa: E402
from tsfmx_b200.tsfm import timesfm as TF  # noqa: E402


def timed(fn, steps, warmup, world, dev):
    for i in range(warmup):
        fn(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    return tdist.allreduce_max(e0.elapsed_time(e1), dev), _lib.launch_count() - launches0


def timesfm_decoder(layers, dev, text_dims=384):
    adapter = TF.TimesFM2p5Adapter(num_layers=layers, precision="bf16", with_quantile_head=False)
    TF.init_random_(adapter, seed=0)
    torch.manual_seed(100)
    return MultimodalDecoder(adapter, MultimodalDecoderConfig(text_dims, 1, [])).to(dev)


def chronos2_decoder(dev, text_dims=384):
    adapter = C2.Chronos2Adapter(precision="bf16")
    C2.init_random_(adapter, seed=0)
    torch.manual_seed(100)
    return MultimodalDecoder(adapter, MultimodalDecoderConfig(text_dims, 1, [])).to(dev)


def batch_for(adapter, b, ctx_len, horizon, seed, dev):
    ctx, masks, text, hor = O.synthetic_batch(b, ctx_len, horizon, seed=seed, patch_len=adapter.patch_len)
    return ctx.to(dev), masks.to(dev), text.to(dev), hor.to(dev)


def emit(rank, name, workload, series, ms, steps, warmup, world, launches, extra=None):
    if rank != 0:
        return
    line = {
        "config": name, "workload": workload, "metric": "series/sec", "value": series * world * steps / (ms * 1e-3),
        "unit": "series/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
        "scaling": "weak", "dtype": "bf16", "data": "synthetic", "gpu_launches": int(launches),
    }
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--finetune-batch", type=int, default=1024, help="series per GPU of the fine-tune step")
    args = ap.parse_args()
    rank, world, local_rank = tdist.init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.check(_lib.load().tsfmx_device_check(local_rank))
    want = lambda n: not args.only or args.only == n  # noqa: E731
    B = args.batch

    if want("chronos2"):
        dec = chronos2_decoder(dev).eval()
        data = [batch_for(dec.adapter, B, 512, 128, 1234 + 17 * rank + i, dev) for i in range(2)]
        with torch.no_grad():
            ms, n = timed(lambda i: dec(128, *data[i % 2][:3]), args.steps, args.warmup, world, dev)
        emit(rank, "cfg3-chronos2", f"Chronos-2 (12 x 768) + 1-layer fusion, ctx 512 / h 128, {B} series per GPU, forecast",
             B, ms, args.steps, args.warmup, world, n)
        del dec, data
        torch.cuda.empty_cache()

    if want("chronos_t5"):
        from tsfmx_b200.tsfm import chronos_t5 as CT5

        tb, horizon = min(B, 2048), 64
        adapter = CT5.ChronosT5Adapter(CT5.ChronosT5Module(), precision="bf16")
        CT5.init_random_(adapter._model, seed=0)
        torch.manual_seed(100)
        dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, [])).to(dev).eval()
        data = []
        for i in range(2):
            ctx, masks, text, _ = O.synthetic_batch(tb, 512, horizon, seed=777 + 17 * rank + i, patch_len=32)
            data.append((ctx.to(dev), masks.to(dev), adapter.expand_text_embeddings(text, 512).to(dev)))
        with torch.no_grad():
            ms, n = timed(lambda i: dec(horizon, *data[i % 2]), max(1, args.steps // 2), 1, world, dev)
            # encoder-only share of the step (tokenise + embed + fusion + 12 encoder blocks)
            def enc_only(i):
                c, m, t = data[i % 2]
                pre = adapter.preprocess(c, m)
                adapter(dec.fusion(pre.input_embeddings, t), pre.masks)
            ms_enc, _ = timed(enc_only, max(1, args.steps // 2), 1, world, dev)
        steps = max(1, args.steps // 2)
        emit(rank, "cfg3-chronos-t5", f"Chronos-T5-base (12 + 12 layers x 768, vocab 4096) + per-token text fusion, ctx 512 "
             f"(513 tokens) / greedy decoding of {horizon} tokens, {tb} series per GPU", tb, ms, steps, 1, world, n,
             {"encoder_ms_per_step": ms_enc / steps})
        del dec, data, adapter
        torch.cuda.empty_cache()

    if want("finetune"):
        fb = min(B, args.finetune_batch)
        dec = timesfm_decoder(50, dev)
        dec.adapter.freeze_parameters()
        dec.train()
        targs = types.SimpleNamespace(per_device_train_batch_size=fb, per_device_eval_batch_size=fb,
                                      gradient_accumulation_steps=1, max_grad_norm=1.0, learning_rate=1e-4,
                                      weight_decay=0.01, num_train_epochs=1, logging_steps=1, seed=0)
        from tsfmx_b200.trainer import MultimodalTrainer

        dummy = [{"context": torch.zeros(512).numpy(), "horizon": torch.zeros(128).numpy(),
                  "text_embeddings": torch.zeros(16, 384).numpy(), "metadata": {}}]
        trainer = MultimodalTrainer(dec, targs, dummy, dummy, "multimodal", dev)
        trainer.rank, trainer.world_size = 0, 1  # the batches below are already this rank's shard
        data = [batch_for(dec.adapter, fb, 512, 128, 4321 + 17 * rank + i, dev) for i in range(2)]

        def step(i):
            c, m, t, h = data[i % 2]
            loss = trainer._forward_loss({"context": c, "horizon": h, "text_embeddings": t})
            loss.backward()
            trainer.optimizer_step()  # all-reduce (NCCL) + clip + AdamW + schedule

        ms, n = timed(step, args.steps, args.warmup, world, dev)
        grads = sum(p.numel() for p in dec.fusion.parameters())
        emit(rank, "cfg4-finetune", f"TimesFM-2.5 layout 50 layers + 1-layer fusion, ctx 512 / h 128, {fb} series per GPU, "
             "fusion fine-tune step (fwd + dgrad + fusion wgrad + all-reduce + AdamW)", fb, ms, args.steps, args.warmup,
             world, n, {"allreduce_bytes_per_step": grads * 4 if world > 1 else 0})
        del dec, trainer, data
        torch.cuda.empty_cache()

    if want("finetune_baseline"):
        fb = min(B, args.finetune_batch)
        dec = timesfm_decoder(50, dev)
        dec.train()
        targs = types.SimpleNamespace(per_device_train_batch_size=fb, per_device_eval_batch_size=fb,
                                      gradient_accumulation_steps=1, max_grad_norm=1.0, learning_rate=1e-5,
                                      weight_decay=0.01, num_train_epochs=1, logging_steps=1, seed=0)
        from tsfmx_b200.trainer import MultimodalTrainer

        dummy = [{"context": torch.zeros(512).numpy(), "horizon": torch.zeros(128).numpy(), "metadata": {}}]
        trainer = MultimodalTrainer(dec, targs, dummy, dummy, "baseline", dev)
        trainer.rank, trainer.world_size = 0, 1  # the batches below are already this rank's shard
        data = [batch_for(dec.adapter, fb, 512, 128, 8321 + 17 * rank + i, dev) for i in range(2)]

        def step_base(i):
            c, m, _t, h = data[i % 2]
            loss = trainer._forward_loss({"context": c, "horizon": h})
            loss.backward()
            trainer.optimizer_step()  # all-reduce of every adapter gradient (NCCL) + clip + AdamW + schedule

        ms, n = timed(step_base, args.steps, args.warmup, world, dev)
        grads = sum(p.numel() for p in dec.adapter.parameters())
        emit(rank, "baseline-full-finetune", f"TimesFM-2.5 layout 50 layers, ctx 512 / h 128, {fb} series per GPU, full "
             "fine-tune step (fwd + dgrad + wgrad of every Linear + all-reduce of all gradients + clip + AdamW)", fb, ms,
             args.steps, args.warmup, world, n, {"allreduce_bytes_per_step": grads * 4 if world > 1 else 0})
        del dec, trainer, data
        torch.cuda.empty_cache()

    if want("finetune_chronos2"):
        fb = min(B, args.finetune_batch)
        dec = chronos2_decoder(dev)
        dec.adapter.freeze_parameters()
        dec.train()
        targs = types.SimpleNamespace(per_device_train_batch_size=fb, per_device_eval_batch_size=fb,
                                      gradient_accumulation_steps=1, max_grad_norm=1.0, learning_rate=1e-4,
                                      weight_decay=0.01, num_train_epochs=1, logging_steps=1, seed=0)
        from tsfmx_b200.trainer import MultimodalTrainer

        dummy = [{"context": torch.zeros(512).numpy(), "horizon": torch.zeros(128).numpy(),
                  "text_embeddings": torch.zeros(32, 384).numpy(), "metadata": {}}]
        trainer = MultimodalTrainer(dec, targs, dummy, dummy, "multimodal", dev)
        trainer.rank, trainer.world_size = 0, 1  # the batches below are already this rank's shard
        data = [batch_for(dec.adapter, fb, 512, 128, 4321 + 17 * rank + i, dev) for i in range(2)]

        def step_c2(i):
            c, m, t, h = data[i % 2]
            loss = trainer._forward_loss({"context": c, "horizon": h, "text_embeddings": t})
            loss.backward()
            trainer.optimizer_step()

        ms, n = timed(step_c2, args.steps, args.warmup, world, dev)
        emit(rank, "cfg4-finetune-chronos2", f"Chronos-2 (12 x 768) + 1-layer fusion, ctx 512 / h 128, {fb} series per GPU, "
             "fusion fine-tune step (fwd + dgrad + fusion wgrad + all-reduce + AdamW)", fb, ms, args.steps, args.warmup,
             world, n)
        del dec, trainer, data
        torch.cuda.empty_cache()

    if want("longctx"):
        dec = timesfm_decoder(50, dev).eval()
        data = [batch_for(dec.adapter, B, 2048, 128, 99 + 17 * rank + i, dev) for i in range(2)]
        with torch.no_grad():
            ms, n = timed(lambda i: dec(128, *data[i % 2][:3]), args.steps, args.warmup, world, dev)
        emit(rank, "cfg5-longctx-timesfm", f"TimesFM-2.5 layout 50 layers + fusion, ctx 2048 / h 128, {B} series per GPU",
             B, ms, args.steps, args.warmup, world, n)
        del dec, data
        torch.cuda.empty_cache()
        dec = chronos2_decoder(dev).eval()
        data = [batch_for(dec.adapter, B, 2048, 256, 199 + 17 * rank + i, dev) for i in range(2)]
        with torch.no_grad():
            ms, n = timed(lambda i: dec(256, *data[i % 2][:3]), args.steps, args.warmup, world, dev)
        emit(rank, "cfg5-longctx-chronos2", f"Chronos-2 (12 x 768) + fusion, ctx 2048 / h 256, {B} series per GPU",
             B, ms, args.steps, args.warmup, world, n)

    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
