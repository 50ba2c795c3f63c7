"""Forecast entry loop on the GPU: staged (copy-stream) batches give the same metrics as the oracle's forecasts."""

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import timesfm_oracle as O  # noqa: E402  (checker only)
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.evaluator import MultimodalEvaluator  # noqa: E402
from tsfmx_b200.tsfm.timesfm import TimesFM2p5Adapter, init_random_  # noqa: E402


def test_evaluator_matches_oracle_metrics():
    adapter = TimesFM2p5Adapter(num_layers=2, with_quantile_head=False)
    init_random_(adapter, seed=0)
    torch.manual_seed(100)
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, []))
    oracle = O.oracle_from_product(dec)
    dec = dec.to("cuda").eval()
    dec.set_precision("bf16x3")
    batches, se, ae, n = [], 0.0, 0.0, 0
    for i, b in enumerate((6, 3, 9, 1)):
        ctx, _m, text, hor = O.synthetic_batch(b, 512, 64, seed=50 + i)
        batches.append({"context": ctx.pin_memory(), "horizon": hor.pin_memory(), "text_embeddings": text.pin_memory(),
                        "metadata": [{}] * b})
        with torch.no_grad():
            p = oracle(64, ctx, torch.zeros_like(ctx, dtype=torch.bool), text)
        se += ((p - hor) ** 2).mean().item() * b
        ae += (p - hor).abs().mean().item() * b
        n += b
    ev = MultimodalEvaluator(dec, torch.device("cuda"))
    got = ev.evaluate(batches)
    assert got["mse"] == pytest.approx(se / n, rel=2e-3) and got["mae"] == pytest.approx(ae / n, rel=2e-3)
    assert ev.evaluate(iter(batches)) == got  # generators work too, and the result is reproducible
    with pytest.raises(RuntimeError, match="empty"):
        ev.evaluate([])


def test_packed_store_feeds_the_evaluator():
    """Pinned zero-copy batches of the packed sample store through the staged evaluator = the per-sample collate path."""
    from tsfmx_b200.data import PackedSamples, multimodal_collate_fn

    adapter = TimesFM2p5Adapter(num_layers=1, with_quantile_head=False)
    init_random_(adapter, seed=0)
    torch.manual_seed(100)
    dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, [])).to("cuda").eval()
    ctx, _m, text, hor = O.synthetic_batch(37, 512, 32, seed=8)
    samples = [{"context": ctx[i].numpy(), "horizon": hor[i].numpy(), "text_embeddings": text[i].numpy(), "metadata": {"i": i}}
               for i in range(37)]
    store = PackedSamples.from_samples(samples)
    assert store.context.is_pinned() and store.text_embeddings.is_pinned()
    ev = MultimodalEvaluator(dec, torch.device("cuda"))
    packed = ev.evaluate(store.batches(8))
    collated = ev.evaluate(multimodal_collate_fn(samples[i : i + 8]) for i in range(0, 37, 8))
    assert packed == collated
    shuffled = ev.evaluate(store.batches(8, shuffle=True, generator=torch.Generator().manual_seed(1)))
    assert shuffled["mse"] == pytest.approx(packed["mse"], rel=1e-5)  # same samples, other batch composition


def _small_decoder(layers=2):
    adapter = TimesFM2p5Adapter(num_layers=layers, with_quantile_head=False)
    init_random_(adapter, seed=0)
    torch.manual_seed(100)
    return MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, [])).to("cuda").eval()


def test_graph_replay_gives_the_eager_forecast():
    """decoder.graphs: the forecast replayed from a CUDA graph is bit-identical to the eager launches, follows the
    contents of the caller's buffers, and is captured afresh when a parameter changes in place."""
    dec = _small_decoder()
    ctx, masks, text, _ = O.synthetic_batch(24, 512, 128, seed=3, padded=True)
    ctx, masks, text = ctx.cuda(), masks.cuda(), text.cuda()
    with torch.no_grad():
        eager = dec.forward_full(128, ctx, masks, text).clone()
        dec.graphs = True
        first = dec.forward_full(128, ctx, masks, text).clone()
        assert torch.equal(first, eager)
        assert len(dec._graph_cache) == 1
        # new contents in the same buffers: same graph, new forecast
        ctx2, masks2, text2, _ = O.synthetic_batch(24, 512, 128, seed=4, padded=True)
        ctx.copy_(ctx2.cuda()); masks.copy_(masks2.cuda()); text.copy_(text2.cuda())
        replayed = dec.forward_full(128, ctx, masks, text).clone()
        assert len(dec._graph_cache) == 1
        dec.graphs = False
        assert torch.equal(replayed, dec.forward_full(128, ctx, masks, text))
        assert not torch.equal(replayed, eager)
        # an in-place parameter update invalidates the graph (the packed bf16 copies are rebuilt)
        dec.graphs = True
        dec.fusion.linears()[0].weight.mul_(1.5)
        updated = dec.forward_full(128, ctx, masks, text).clone()
        dec.graphs = False
        assert torch.equal(updated, dec.forward_full(128, ctx, masks, text))
        assert not torch.equal(updated, replayed)
        # no text embeddings, other horizon: separate graphs
        dec.graphs = True
        assert torch.equal(dec.forward_full(64, ctx, masks), dec._forecast(64, ctx, masks, None))
    assert len(dec._graph_cache) <= MultimodalDecoder.GRAPH_CACHE_ENTRIES


def test_evaluator_graphs_match_eager_and_are_reused():
    dec = _small_decoder()
    batches = []
    for i, b in enumerate((16, 16, 16, 5)):  # ragged last batch
        ctx, _m, text, hor = O.synthetic_batch(b, 512, 64, seed=70 + i)
        batches.append({"context": ctx.pin_memory(), "horizon": hor.pin_memory(), "text_embeddings": text.pin_memory()})
    eager = MultimodalEvaluator(dec, torch.device("cuda"), graphs=False).evaluate(batches)
    ev = MultimodalEvaluator(dec, torch.device("cuda"))
    assert ev.graphs and not dec.graphs
    got = ev.evaluate(batches)
    assert got == eager
    assert not dec.graphs  # restored
    captured = len(dec._graph_cache)
    assert 1 <= captured <= 3  # two staging slots, plus the ragged tail of one of them
    keys = set(dec._graph_cache)
    assert ev.evaluate(batches) == eager  # second pass: stable staging addresses, nothing captured again
    assert set(dec._graph_cache) == keys


def test_evaluator_stops_capturing_when_shapes_keep_changing():
    dec = _small_decoder(1)
    batches = []
    for i, b in enumerate((9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20)):  # a new batch size every time
        ctx, _m, text, hor = O.synthetic_batch(b, 512, 32, seed=90 + i)
        batches.append({"context": ctx.pin_memory(), "horizon": hor.pin_memory(), "text_embeddings": text.pin_memory()})
    eager = MultimodalEvaluator(dec, torch.device("cuda"), graphs=False).evaluate(batches)
    ev = MultimodalEvaluator(dec, torch.device("cuda"))
    assert ev.evaluate(batches) == eager
    assert dec.graph_captures == MultimodalEvaluator.MAX_CAPTURES_PER_PASS
    assert not dec.graphs


def test_shuffled_store_with_a_lagging_copy_stream():
    """Large shuffled batches while the copy stream runs far behind the host (a spin kernel parked on it): every sample
    must still be evaluated exactly once.  With one reusable pinned gather buffer the host overwrote batch i while its
    DMA was still queued, and the metrics came out silently wrong."""
    from tsfmx_b200.data import PackedSamples

    dec = _small_decoder(1)
    n, b = 1536, 256  # 6 batches of 256 series: 6.3 MB of text embeddings per batch
    ctx, _m, text, hor = O.synthetic_batch(n, 512, 32, seed=21)
    hor = hor + torch.arange(n, dtype=torch.float32)[:, None] * 0.01  # per-sample targets: duplicates shift the metrics
    store = PackedSamples(ctx.pin_memory(), hor.pin_memory(), text.pin_memory(), [{"i": i} for i in range(n)])
    ev = MultimodalEvaluator(dec, torch.device("cuda"))
    in_order = ev.evaluate(store.batches(b))
    for seed in (1, 2):
        ev._copy_stream = ev._copy_stream or torch.cuda.Stream()
        with torch.cuda.stream(ev._copy_stream):
            torch.cuda._sleep(400_000_000)  # ~0.2 s: the host gathers all six batches before the first DMA starts
        got = ev.evaluate(store.batches(b, shuffle=True, generator=torch.Generator().manual_seed(seed)))
        assert got["mse"] == pytest.approx(in_order["mse"], rel=1e-5)
        assert got["mae"] == pytest.approx(in_order["mae"], rel=1e-5)


@pytest.mark.parametrize("graphs", [False, True])
def test_predict_returns_every_forecast_on_the_host(graphs):
    """MultimodalEvaluator.predict: staged H2D like evaluate, every forecast copied back on its own stream while the
    next batch computes.  Forecasts equal direct forward_full calls, batch order and ragged tail preserved; copy=False
    hands out views of the page-locked ring."""
    dec = _small_decoder()
    batches, want = [], []
    for i, b in enumerate((12, 12, 12, 12, 12, 5)):
        ctx, _m, text, hor = O.synthetic_batch(b, 512, 64, seed=40 + i)
        batches.append({"context": ctx.pin_memory(), "horizon": hor.pin_memory(), "text_embeddings": text.pin_memory()})
        with torch.no_grad():
            want.append(dec.forward_full(64, ctx.cuda(), torch.zeros_like(ctx, dtype=torch.bool).cuda(), text.cuda()).cpu())
    ev = MultimodalEvaluator(dec, torch.device("cuda"), graphs=graphs)
    got = list(ev.predict(batches))
    assert [g.shape for g in got] == [w.shape for w in want]
    for g, w in zip(got, want):
        assert not g.is_cuda and torch.equal(g, w)
    points = list(ev.predict(iter(batches), horizon=32, full=False))
    assert points[0].shape == (12, 32) and points[-1].shape == (5, 32)
    with torch.no_grad():
        ctx = batches[3]["context"]
        ref = dec(32, ctx.cuda(), torch.zeros_like(ctx, dtype=torch.bool).cuda(), batches[3]["text_embeddings"].cuda()).cpu()
    assert torch.equal(points[3], ref)
    seen = 0
    for g, w in zip(ev.predict(batches, copy=False), want):  # views: consumed before two more batches are requested
        assert g.is_pinned() and torch.equal(g, w)
        seen += 1
    assert seen == len(want)
    assert list(ev.predict([])) == []


def test_graph_survives_a_precision_round_trip():
    """set_precision("bf16x3") and back re-packs the weights into new buffers; the bf16 graph captured before has the
    same cache key again and must still read the (kept alive) weights it was captured with."""
    dec = _small_decoder()
    ctx, masks, text, _ = O.synthetic_batch(24, 512, 128, seed=5)
    ctx, masks, text = ctx.cuda(), masks.cuda(), text.cuda()
    with torch.no_grad():
        dec.set_precision("bf16")
        eager = dec.forward_full(128, ctx, masks, text).clone()
        dec.graphs = True
        assert torch.equal(dec.forward_full(128, ctx, masks, text), eager)
        dec.set_precision("bf16x3")
        precise = dec.forward_full(128, ctx, masks, text).clone()
        junk = [torch.randn(64, 1280, 1280, device="cuda") for _ in range(4)]  # churn the allocator over freed blocks
        dec.set_precision("bf16")
        again = dec.forward_full(128, ctx, masks, text).clone()
        dec.graphs = False
    del junk
    assert torch.equal(again, eager) and not torch.equal(precise, eager)
