"""torchrun check of the data-parallel fusion fine-tune step (NCCL all-reduce over NVLink):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/ddp_finetune_check.py

Every rank trains the same seeded model on its shard of each batch with MultimodalTrainer; after the all-reduce the
fusion gradients (and, after the optimizer step, the weights) must equal the single-process global-batch result that
every rank also computes locally, and all ranks must hold identical weights.
"""

import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "multimodal-timesfm_b200"))
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import timesfm_oracle as O  # noqa: E402  (synthetic batch generator)
from tsfmx_b200 import distributed as tdist  # noqa: E402
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.trainer import MultimodalTrainer  # noqa: E402
from tsfmx_b200.tsfm.timesfm import TimesFM2p5Adapter, init_random_  # noqa: E402


def build(device):
    adapter = TimesFM2p5Adapter(num_layers=2, with_quantile_head=False)
    init_random_(adapter, seed=0)
    torch.manual_seed(100)
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, [])).to(device)
    dec.set_precision("bf16x3")
    return dec


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "multimodal"   # or "baseline": full fine-tuning of the adapter
    rank, world, local_rank = tdist.init_process_group("nccl")
    device = torch.device("cuda", local_rank)
    per_rank, steps = 4, 3
    ctx, masks, text, hor = O.synthetic_batch(per_rank * world * steps, 512, 64, seed=77)
    samples = [
        {"context": ctx[i].numpy(), "horizon": hor[i].numpy(), "text_embeddings": text[i].numpy(), "metadata": {"i": i}}
        for i in range(ctx.shape[0])
    ]
    args = types.SimpleNamespace(per_device_train_batch_size=per_rank, per_device_eval_batch_size=per_rank,
                                 gradient_accumulation_steps=1, max_grad_norm=1.0, learning_rate=1e-3, weight_decay=0.01,
                                 num_train_epochs=1, logging_steps=1, seed=0)
    # distributed run: each rank sees its slice of every global batch
    dec = build(device)
    trainer = MultimodalTrainer(dec, args, samples, samples[: per_rank * world], mode, device)
    loss = trainer.train_epoch()
    pick = (lambda d: d.fusion.linears()[0].weight) if mode == "multimodal" else (
        lambda d: d.adapter._model.stacked_xf[0].ff0.weight)
    w_dist = pick(dec).detach().clone()

    # single-process reference on the same global batches (world size forced to 1 for this trainer)
    ref = build(device)
    ref_trainer = MultimodalTrainer(ref, args, samples, samples[: per_rank * world], mode, device)
    ref_trainer.rank, ref_trainer.world_size = 0, 1   # no sharding, no collectives
    ref.grad_ready_hook = None
    ref_loss = ref_trainer.train_epoch()
    w_ref = pick(ref).detach()

    rel = ((w_dist - w_ref).norm() / (w_ref - pick(build(device))).norm()).item()
    same = True
    if world > 1:
        gathered = [torch.empty_like(w_dist) for _ in range(world)]
        dist.all_gather(gathered, w_dist)
        same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        overlapped = dec.grad_ready_hook is not None
        print(f"mode={mode} world={world} steps={trainer.global_step} overlapped_allreduce={overlapped} "
              f"graph_replays={trainer.graph_replays} loss dist={loss:.6f} ref={ref_loss:.6f} "
              f"update rel diff={rel:.3e} identical_across_ranks={same}")
    assert same, "ranks diverged"
    assert abs(loss - ref_loss) < 1e-4 * max(1.0, abs(ref_loss)), (loss, ref_loss)
    assert rel < 2e-3, rel
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print("ddp fine-tune check ok")


if __name__ == "__main__":
    main()
