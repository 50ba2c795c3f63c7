"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: series sharding, the single flattened gradient
all-reduce of the fine-tune step, and the bench-style max-over-ranks reduction."""

import os
import socket

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
