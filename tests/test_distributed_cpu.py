"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: series sharding, the single flattened gradient
all-reduce of the fine-tune step, and the bench-style max-over-ranks reduction."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tsfmx_b200 import distributed as tdist


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions_the_batch():
    for total in (0, 1, 7, 8, 4096, 16385):
        for world in (1, 2, 3, 8):
            spans = [tdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (lo0, hi0), (lo1, _hi1) in zip(spans, spans[1:]):
                assert hi0 == lo1 and hi0 >= lo0
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank: int, world: int, port: int, result_dir: str):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = tdist.init_process_group("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    # data-parallel equivalence: mean-over-shard loss + mean all-reduce == single-process global-batch gradient
    fusion = torch.nn.Linear(12, 5, bias=False)
    x, y = torch.randn(8, 12), torch.randn(8, 5)
    full = torch.nn.functional.mse_loss(torch.relu(fusion(x)), y)
    (g_full,) = torch.autograd.grad(full, fusion.weight)
    batch = tdist.shard_batch({"context": x, "horizon": y, "metadata": list(range(8))}, rank, world)
    assert len(batch["metadata"]) == 4 and batch["metadata"][0] == 4 * rank
    loss = torch.nn.functional.mse_loss(torch.relu(fusion(batch["context"])), batch["horizon"])
    loss.backward()
    extra = torch.full((3,), float(rank))
    tdist.allreduce_mean_([fusion.weight.grad, None, extra])
    assert torch.allclose(fusion.weight.grad, g_full, atol=1e-6)
    assert torch.allclose(extra, torch.full((3,), (world - 1) / 2))
    assert tdist.allreduce_max(10.0 + rank, torch.device("cpu")) == 10.0 + world - 1
    # bucketed path of the full fine-tune (all adapter gradients): several flattened collectives, a tensor larger than
    # a bucket on its own, mixed shapes; every element must come back as the mean over ranks
    saved, tdist.BUCKET_ELEMS = tdist.BUCKET_ELEMS, 64
    try:
        gen = torch.Generator().manual_seed(5)
        shapes = [(7, 5), (64,), (3, 3, 3), (200,), (1,), (63,), (2, 40)]
        base = [torch.randn(sh, generator=gen) for sh in shapes]
        mine = [b * (rank + 1) for b in base]
        tdist.allreduce_mean_(mine)
        scale = sum(r + 1 for r in range(world)) / world
        for got, b in zip(mine, base):
            assert torch.allclose(got, b * scale, atol=1e-6)
    finally:
        tdist.BUCKET_ELEMS = saved
    # overlapped reducer of the full fine-tune: groups of gradients handed over layer by layer, large ones reduced in
    # place as they arrive, small ones in one flattened bucket at finish()
    saved_direct, tdist.DIRECT_ELEMS = tdist.DIRECT_ELEMS, 32
    try:
        reducer = tdist.OverlappedGradReducer()
        gen = torch.Generator().manual_seed(9)
        groups = [[torch.randn(8, 8, generator=gen), torch.randn(5, generator=gen)], [torch.randn(3, 40, generator=gen)],
                  [torch.randn(7, generator=gen), None, torch.randn(6, 6, generator=gen)]]
        mine = [[None if t is None else t * (rank + 1) for t in g] for g in groups]
        for g in mine:
            reducer(g)
        assert not reducer.reduced
        reducer.finish()
        assert reducer.reduced and reducer.bytes == 4 * sum(t.numel() for g in groups for t in g if t is not None)
        total = sum(r + 1 for r in range(world))
        for g_mine, g_base in zip(mine, groups):
            for got, b in zip(g_mine, g_base):
                if b is not None:
                    assert torch.allclose(got, b * total, atol=1e-5)
    finally:
        tdist.DIRECT_ELEMS = saved_direct
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(result_dir, f"ok{rank}"), "w").close()


def test_two_rank_gloo_allreduce(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["ok0", "ok1"]


def test_single_process_is_a_no_op():
    t = torch.ones(4)
    tdist.allreduce_mean_([t])
    assert torch.equal(t, torch.ones(4))
    assert tdist.world_info()[1] >= 1


# ------------------------------------------------------------------------------------------------ trainer, ragged tail
class _StubAdapter(torch.nn.Module):
    """Stands in for a frozen backbone so that the trainer's HOST logic (sharding, loss weighting, the gradient
    all-reduce, the optimizer step order) can run on CPU ranks; the product adapters are CUDA-only."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.ones(1))

    def freeze_parameters(self):
        self.w.requires_grad = False


class _StubDecoder(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.adapter = _StubAdapter()
        self.fusion = torch.nn.Linear(6, 4, bias=False)

    def forward(self, horizon, context, padding, text):
        return context[:, -horizon:] * self.adapter.w + torch.relu(self.fusion(text.mean(1)))[:, :horizon]


def _trainer_dataset(n=5):
    g = torch.Generator().manual_seed(3)
    return [{"context": torch.randn(8, generator=g).numpy(), "horizon": torch.randn(4, generator=g).numpy(),
             "text_embeddings": torch.randn(2, 6, generator=g).numpy(), "metadata": {"i": i}} for i in range(n)]


def _run_trainer(per_device: int, epochs: int = 2):
    import types

    from tsfmx_b200.trainer import MultimodalTrainer

    torch.manual_seed(11)
    model = _StubDecoder()
    args = types.SimpleNamespace(per_device_train_batch_size=per_device, per_device_eval_batch_size=per_device,
                                 gradient_accumulation_steps=1, max_grad_norm=0.5, learning_rate=1e-2, weight_decay=0.0,
                                 num_train_epochs=epochs, seed=4, warmup_steps=0.0, save_strategy="no")
    data = _trainer_dataset()
    tr = MultimodalTrainer(model, args, data, data, "multimodal", torch.device("cpu"))
    losses = [tr.train_epoch() for _ in range(epochs)]
    return model.fusion.weight.detach().clone(), losses, tr.validate_epoch()


def _trainer_worker(rank: int, world: int, port: int, result_dir: str):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    tdist.init_process_group("gloo")
    w, losses, val = _run_trainer(per_device=2)  # global batch 4 over 5 samples: the tail batch has ONE sample
    torch.save({"w": w, "losses": losses, "val": val}, os.path.join(result_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_trainer_ragged_tail_equals_the_single_process_step(tmp_path):
    """5 samples, 2 ranks x 2 per device: the last global batch holds one sample, so rank 1's shard is EMPTY.  Every
    rank must still join the gradient all-reduce (no hang, no NaN), and weights and losses must equal the
    single-process run over the same global batches (4 + 1)."""
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    ref_w, ref_losses, ref_val = _run_trainer(per_device=4)
    world, port = 2, _free_port()
    mp.spawn(_trainer_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        os.environ.pop(k, None)
    outs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert torch.equal(outs[0]["w"], outs[1]["w"])
    assert torch.allclose(outs[0]["w"], ref_w, atol=1e-6)
    for got in outs:
        assert got["losses"] == pytest.approx(ref_losses, rel=1e-5)
        assert got["val"] == pytest.approx(ref_val, rel=1e-5)
        assert all(v == v for v in got["losses"])  # no NaN from the empty shard
