"""The product's MultimodalTrainer against the REFERENCE's own MultimodalTrainer (imported from /root/reference/src, build
container only - the GPU box has no /root/reference and skips): both drive copies of the same CPU oracle decoder over
the same data, so everything between the batch and the weights - loss, accumulation, clip, AdamW, the warm-up schedule
stepped per optimizer step, the epoch loss - is compared with the reference's real code (trainer.py:186-283), not with
a restatement.  Every batch is the whole dataset, so the two trainers' different shuffling seeds cannot matter."""

import copy
import sys
from pathlib import Path

import pytest
import torch

from oracle import timesfm_oracle as O
from tsfmx_b200.trainer import MultimodalTrainer

REF_SRC = Path("/root/reference/src")


def _reference_classes():
    if not REF_SRC.exists():
        pytest.skip("the reference tree is only present in the build container")
    sys.path.insert(0, str(REF_SRC))
    try:
        from tsfmx.trainer import MultimodalTrainer as RefTrainer
        from tsfmx.training_args import TrainingArguments
    finally:
        sys.path.remove(str(REF_SRC))
    return RefTrainer, TrainingArguments


def _samples(n, seed):
    ctx, _m, text, hor = O.synthetic_batch(n, 128, 32, seed=seed)
    return [{"context": ctx[i].numpy(), "horizon": hor[i].numpy(), "text_embeddings": text[i].numpy(), "metadata": {"i": i}}
            for i in range(n)]


@pytest.mark.parametrize("mode,accum", [("multimodal", 1), ("baseline", 1), ("multimodal", 2)])
def test_trainer_equals_the_reference_trainer(tmp_path, mode, accum):
    RefTrainer, TrainingArguments = _reference_classes()
    torch.manual_seed(0)
    adapter = O.OracleTimesFM2p5Adapter(1)
    for p in adapter.parameters():
        torch.nn.init.normal_(p, std=0.05)
    model_ref = O.OracleDecoder(adapter, 384, 1, [])
    model_ref.train()
    model_mine = copy.deepcopy(model_ref)
    untouched = copy.deepcopy(model_ref)
    train, val = _samples(8, 5), _samples(4, 6)
    # one batch = the whole dataset (two half-batches under accumulation would depend on the shuffle: accum = 2 uses a
    # dataset of two identical halves, so that any split gives the same two micro-batches up to row order)
    if accum == 2:
        train = train[:4] + train[:4]
    args = TrainingArguments(output_dir=str(tmp_path / "out"), per_device_train_batch_size=len(train) // accum,
                             per_device_eval_batch_size=4, num_train_epochs=6, learning_rate=3e-3, weight_decay=0.01,
                             warmup_steps=0.34, gradient_accumulation_steps=accum, max_grad_norm=0.5,
                             lr_scheduler_type="cosine", logging_strategy="no", eval_strategy="no", save_strategy="no", seed=1)
    ref = RefTrainer(model_ref, args, train, val, mode, torch.device("cpu"), None)
    mine = MultimodalTrainer(model_mine, args, train, val, mode, torch.device("cpu"))
    assert not mine.graphs
    if accum == 2:  # two micro-batches per step: both loaders are rebuilt unshuffled so that they see the same halves
        from torch.utils.data import DataLoader
        for t in (ref, mine):
            t.train_loader = DataLoader(train, batch_size=args.per_device_train_batch_size, shuffle=False, num_workers=0,
                                        collate_fn=t.train_loader.collate_fn)
    for epoch in range(args.num_train_epochs):
        loss_ref = ref.train_epoch()
        loss_mine = mine.train_epoch()
        assert loss_mine == pytest.approx(loss_ref, rel=2e-5), (epoch, loss_mine, loss_ref)
        assert mine.global_step == ref.global_step == epoch + 1
        assert mine.optimizer.param_groups[0]["lr"] == pytest.approx(ref.optimizer.param_groups[0]["lr"], rel=1e-12)
    assert mine.validate_epoch() == pytest.approx(ref.validate_epoch(), rel=2e-5)
    params_ref = dict(model_ref.named_parameters())
    changed = 0
    for name, p in model_mine.named_parameters():
        q = params_ref[name]
        assert ((p - q).norm() / q.norm().clamp_min(1e-30)).item() < 1e-5, name  # batch row order differs: fp32 rounding
        changed += int(not torch.equal(p, dict(untouched.named_parameters())[name]))
    assert changed == (1 if mode == "multimodal" else len(list(adapter.parameters())))  # exactly the trained tensors moved
