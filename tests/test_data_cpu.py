"""Packed sample store against the reference's collate path (reference tsfmx/data/collate.py:9-29): same batches."""

import pickle

import numpy as np
import pytest
import torch

from tsfmx_b200.data import PackedSamples, baseline_collate_fn, multimodal_collate_fn


def _samples(n, with_text=True, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        s = {"context": rng.standard_normal(64).astype(np.float32), "horizon": rng.standard_normal(16).astype(np.float32),
             "metadata": {"i": i}}
        if with_text:
            s["text_embeddings"] = rng.standard_normal((2, 8)).astype(np.float32)
        out.append(s)
    return out


@pytest.mark.parametrize("with_text", [True, False])
def test_in_order_batches_equal_collate(with_text, tmp_path):
    samples = _samples(11, with_text)
    path = tmp_path / "cache.pkl"
    with open(path, "wb") as f:
        pickle.dump(samples, f)                       # the reference's cache format: pickled list of sample dicts
    store = PackedSamples.from_pickle(path, pin=False)
    assert len(store) == 11 and np.array_equal(store[3]["context"], samples[3]["context"])
    collate = multimodal_collate_fn if with_text else baseline_collate_fn
    got = list(store.batches(4))
    assert [len(b["context"]) for b in got] == [4, 4, 3]
    for k, b in enumerate(got):
        ref = collate(samples[4 * k : 4 * k + 4])
        assert set(b) == set(ref)
        for key in ref:
            if key == "metadata":
                assert b[key] == ref[key]
            else:
                assert torch.equal(b[key], ref[key])
        assert b["context"].data_ptr() == store.context[4 * k].data_ptr()   # zero-copy slice
    assert [len(b["context"]) for b in store.batches(4, drop_last=True)] == [4, 4]


def test_shuffled_epoch_is_a_permutation_and_reproducible():
    samples = _samples(10)
    store = PackedSamples.from_samples(samples, pin=False)
    seen = []
    for b in store.batches(3, shuffle=True, generator=torch.Generator().manual_seed(5)):
        ids = [m["i"] for m in b["metadata"]]
        for row, i in enumerate(ids):
            assert np.array_equal(b["context"][row].numpy(), samples[i]["context"])
            assert np.array_equal(b["text_embeddings"][row].numpy(), samples[i]["text_embeddings"])
        seen += ids
    assert sorted(seen) == list(range(10))
    again = [m["i"] for b in store.batches(3, shuffle=True, generator=torch.Generator().manual_seed(5)) for m in b["metadata"]]
    assert again == seen


def test_rejects_empty_and_ragged():
    with pytest.raises(ValueError, match="empty"):
        PackedSamples.from_samples([])
    bad = _samples(3)
    bad[2]["context"] = np.zeros(32, dtype=np.float32)
    with pytest.raises(ValueError, match="ragged"):
        PackedSamples.from_samples(bad, pin=False)
    with pytest.raises(ValueError, match="batch_size"):
        list(PackedSamples.from_samples(_samples(2), pin=False).batches(0))


def test_shuffled_batches_do_not_alias_each_other():
    """A consumer may hold (or still be copying) earlier shuffled batches when it asks for the next one: every batch
    owns its storage.  (A shared scratch buffer here silently corrupted evaluations whose H2D copies ran behind.)"""
    samples = _samples(12)
    store = PackedSamples.from_samples(samples, pin=False)
    held = list(store.batches(4, shuffle=True, generator=torch.Generator().manual_seed(9)))
    assert len({b["context"].data_ptr() for b in held}) == len(held)
    for b in held:
        for row, meta in enumerate(b["metadata"]):
            assert np.array_equal(b["context"][row].numpy(), samples[meta["i"]]["context"])
            assert np.array_equal(b["horizon"][row].numpy(), samples[meta["i"]]["horizon"])
            assert np.array_equal(b["text_embeddings"][row].numpy(), samples[meta["i"]]["text_embeddings"])
