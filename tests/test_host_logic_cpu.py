"""Host-side logic that needs no GPU: series-lane helpers, the decoder's error behaviour before any device work
(same ``ValueError``s as the reference, reference tsfmx/decoder.py:62-63, fusion.py:36-42, timesfm.py:48-51,116-119),
and the loud failure of every adapter on CPU tensors (there is no CPU fallback)."""

import pytest
import torch

from tsfmx_b200 import lanes
from tsfmx_b200._lib import TsfmxError
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig
from tsfmx_b200.fusion import MultimodalFusion
from tsfmx_b200.tsfm.chronos import Chronos2Adapter, Chronos2Module
from tsfmx_b200.tsfm.chronos_t5 import ChronosT5Adapter, ChronosT5Module
from tsfmx_b200.tsfm.timesfm import TimesFM2p5Adapter


def test_split_points_cover_the_batch_once():
    for batch, count, mult in [(4096, 2, 8), (1031, 2, 8), (10, 4, 1), (3, 8, 8), (1, 2, 8)]:
        cuts = lanes.split_points(batch, count, mult)
        assert cuts[0][0] == 0 and cuts[-1][1] == batch
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        assert all(hi > lo for lo, hi in cuts) and len(cuts) <= max(1, min(count, batch))


def test_step_generators_drain_and_fall_back():
    class Stage:
        def plain(self, x):
            return x + 1

        def fancy_steps(self, x):
            yield
            yield
            return x * 2

    s = Stage()
    assert lanes.drain(lanes.steps(s, "plain", 3)) == 4      # no *_steps variant: one step
    assert lanes.drain(lanes.steps(s, "fancy", 3)) == 6      # generator's return value comes through
    assert len(list(lanes.steps(s, "fancy", 3))) == 2


def test_lane_count_needs_enough_tokens_per_lane():
    dec = MultimodalDecoder(TimesFM2p5Adapter(num_layers=1, with_quantile_head=False), MultimodalDecoderConfig())
    assert dec._lane_count(torch.zeros(4096, 512)) == 1          # CPU tensors never split
    dec.lanes = 1
    assert dec._lane_count(torch.zeros(4096, 512)) == 1


def test_reference_error_behaviour_before_device_work():
    dec = MultimodalDecoder(TimesFM2p5Adapter(num_layers=1, with_quantile_head=False), MultimodalDecoderConfig())
    with pytest.raises(ValueError, match="must match inputs shape"):
        dec.forward_full(128, torch.zeros(2, 512), torch.zeros(2, 256, dtype=torch.bool))
    with pytest.raises(ValueError, match="divisible by patch length"):
        dec.adapter.preprocess(torch.zeros(2, 500), torch.zeros(2, 500, dtype=torch.bool))
    with pytest.raises(ValueError, match="horizon must be <= output_patch_len"):
        dec.adapter.postprocess(129, torch.zeros(2, 16, 1280), {})
    with pytest.raises(ValueError, match="between 1 and 3"):
        MultimodalFusion(1280, 384, num_layers=4, hidden_dims=[8, 8, 8])
    with pytest.raises(ValueError, match="hidden_dims must have"):
        MultimodalFusion(1280, 384, num_layers=2, hidden_dims=[])
    with pytest.raises(ValueError, match="exceeds the maximum prediction length"):
        Chronos2Adapter(Chronos2Module(1)).postprocess(1025, torch.zeros(1, 64, 768), {})


def test_adapters_refuse_cpu_tensors():
    x, m = torch.zeros(2, 512), torch.zeros(2, 512, dtype=torch.bool)
    for adapter in (TimesFM2p5Adapter(num_layers=1, with_quantile_head=False), Chronos2Adapter(Chronos2Module(1)),
                    ChronosT5Adapter(ChronosT5Module(num_layers=1))):
        with pytest.raises(TsfmxError, match="no CPU fallback"):
            adapter.preprocess(x, m)
    with pytest.raises(TsfmxError, match="no CPU fallback"):
        MultimodalFusion(1280, 384).forward_device(torch.zeros(2, 16, 1280), torch.zeros(2, 16, 384))


def test_state_dict_keys_follow_upstream_names():
    tf = TimesFM2p5Adapter(num_layers=2, with_quantile_head=True)._model.state_dict()
    for k in ("tokenizer.hidden_layer.weight", "stacked_xf.1.attn.qkv_proj.weight", "stacked_xf.0.attn.per_dim_scale.per_dim_scale",
              "stacked_xf.0.pre_attn_ln.scale", "output_projection_point.residual_layer.weight",
              "output_projection_quantiles.output_layer.weight"):
        assert k in tf, k
    t5 = ChronosT5Adapter(ChronosT5Module(num_layers=1))._model.state_dict()
    for k in ("shared.weight", "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight",
              "decoder.block.0.layer.1.EncDecAttention.k.weight", "decoder.block.0.layer.2.DenseReluDense.wo.weight",
              "encoder.final_layer_norm.weight"):
        assert k in t5, k
    dec = MultimodalDecoder(TimesFM2p5Adapter(num_layers=1, with_quantile_head=False), MultimodalDecoderConfig(384, 3, [512, 256]))
    assert [k for k in dec.fusion.state_dict()] == ["projection.0.weight", "projection.2.weight", "projection.4.weight"]


def test_trainer_checkpoint_strategies(tmp_path):
    """save_checkpoint / rotation / best-model reload follow the reference (trainer.py:285-354); pure host logic, driven
    on a trainer object whose model never runs."""
    import types

    from tsfmx_b200.trainer import MultimodalTrainer

    dec = MultimodalDecoder(TimesFM2p5Adapter(num_layers=1, with_quantile_head=False), MultimodalDecoderConfig())
    args = types.SimpleNamespace(per_device_train_batch_size=2, per_device_eval_batch_size=2, gradient_accumulation_steps=1,
                                 max_grad_norm=1.0, learning_rate=1e-3, weight_decay=0.0, num_train_epochs=1, seed=0,
                                 save_strategy="epoch", save_total_limit=2, checkpoint_dir=tmp_path / "ckpt")
    dummy = [{"context": torch.zeros(64).numpy(), "horizon": torch.zeros(8).numpy(),
              "text_embeddings": torch.zeros(2, 384).numpy(), "metadata": {}}] * 2
    tr = MultimodalTrainer(dec, args, dummy, dummy, "multimodal", torch.device("cpu"))
    assert all(not p.requires_grad for p in dec.adapter.parameters())  # multimodal mode freezes the adapter
    for epoch, val in enumerate([0.9, 0.5, 0.7, 0.6]):
        tr.current_epoch = epoch
        tr.save_checkpoint(val)
    names = sorted(p.name for p in args.checkpoint_dir.iterdir())
    assert names == ["best_model.pt", "checkpoint_epoch_2.pt", "checkpoint_epoch_3.pt"]
    best = torch.load(args.checkpoint_dir / "best_model.pt", weights_only=True)
    assert best["epoch"] == 1 and best["best_val_loss"] == 0.5
    assert set(best) == {"epoch", "global_step", "optimizer_state_dict", "scheduler_state_dict", "best_val_loss",
                         "fusion_state_dict"}
    with torch.no_grad():
        dec.fusion.linears()[0].weight.zero_()
    tr._load_checkpoint_state(best)
    assert float(dec.fusion.linears()[0].weight.detach().abs().sum()) > 0
    args.save_strategy = "best"
    tr.current_epoch = 4
    tr.save_checkpoint(0.8)  # not an improvement: nothing written
    assert not (args.checkpoint_dir / "checkpoint_epoch_4.pt").exists()
    args.eval_strategy = "steps"
    with pytest.raises(NotImplementedError):
        tr.train()
    args.eval_strategy = "epoch"
    base = MultimodalTrainer(dec, args, dummy, dummy, "baseline", torch.device("cpu"))   # TimesFM: full fine-tuning
    assert all(p.requires_grad for p in dec.adapter.parameters())
    assert "adapter_state_dict" in base.build_checkpoint() and "fusion_state_dict" not in base.build_checkpoint()
    assert len(list(base._get_trainable_params())) == len(list(dec.adapter.parameters()))
    c2 = MultimodalDecoder(Chronos2Adapter(Chronos2Module(1)), MultimodalDecoderConfig())
    MultimodalTrainer(c2, args, dummy, dummy, "baseline", torch.device("cpu"))   # Chronos-2: full fine-tuning too
    assert all(p.requires_grad for p in c2.adapter.parameters())
    t5 = MultimodalDecoder(ChronosT5Adapter(ChronosT5Module(num_layers=1)), MultimodalDecoderConfig())
    with pytest.raises(NotImplementedError):
        MultimodalTrainer(t5, args, dummy, dummy, "baseline", torch.device("cpu"))
    with pytest.raises(ValueError):
        MultimodalTrainer(dec, args, dummy, dummy, "something", torch.device("cpu"))


def test_warmup_steps_follow_the_reference_semantics():
    """``TrainingArguments.warmup_steps`` (reference training_args.py:28-35,111-121): a float; >= 1 means absolute steps,
    [0, 1) a ratio of the total, rounded UP.  ``warmup_steps=0.1`` must not truncate to zero warm-up steps."""
    import math
    import types

    from tsfmx_b200.trainer import MultimodalTrainer

    dec = MultimodalDecoder(TimesFM2p5Adapter(num_layers=1, with_quantile_head=False), MultimodalDecoderConfig())
    dummy = [{"context": torch.zeros(64).numpy(), "horizon": torch.zeros(8).numpy(),
              "text_embeddings": torch.zeros(2, 384).numpy(), "metadata": {}}] * 6

    def trainer(**kw):
        args = types.SimpleNamespace(per_device_train_batch_size=2, per_device_eval_batch_size=2, gradient_accumulation_steps=1,
                                     max_grad_norm=1.0, learning_rate=1.0, weight_decay=0.0, num_train_epochs=7, seed=0, **kw)
        return MultimodalTrainer(dec, args, dummy, dummy, "multimodal", torch.device("cpu"))

    total = 7 * 3
    for ws, want in [(0.0, 0), (0.1, math.ceil(total * 0.1)), (0.5, 11), (1.0, 1), (5, 5), (5.9, 5)]:
        tr = trainer(warmup_steps=ws)
        assert tr._warmup_steps(total) == want, (ws, want)
    tr = trainer(warmup_steps=0.1)  # 3 warm-up steps: the LR ramps 0, 1/3, 2/3, 1 instead of starting at 1
    assert tr.optimizer.param_groups[0]["lr"] == 0.0
    # an args object that carries the reference's own method is asked directly
    tr = trainer(warmup_steps=0.25, get_warmup_steps=lambda n: 4)
    assert tr._warmup_steps(total) == 4
    try:
        import sys
        sys.path.insert(0, "/root/reference/src")
        from tsfmx.training_args import TrainingArguments  # only in the build container
    except Exception:
        return
    finally:
        if sys.path[0] == "/root/reference/src":
            sys.path.pop(0)
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        for ws in (0.0, 0.1, 0.5, 1.0, 5.0):
            ref = TrainingArguments(output_dir=d, warmup_steps=ws)
            assert trainer(warmup_steps=ws)._warmup_steps(total) == ref.get_warmup_steps(total)


def test_wgrad_takes_token_major_operands_only_when_they_qualify():
    """`ops.wgrad` hands bf16 row-major operands (16-byte aligned rows) to the token-major GEMM; split, fp32, strided-by-odd
    or narrow views go through the transposing path.  (Host-side dispatch only: no kernel runs here.)"""
    import torch

    from tsfmx_b200 import ops

    x = torch.zeros(64, 1280, dtype=torch.bfloat16)
    assert ops._token_major_ok(x, 1280)
    assert ops._token_major_ok(x[:, 8:72], 64)  # aligned column window of a wider matrix
    assert not ops._token_major_ok(x[:, 4:68], 64)  # starts 8 bytes into a row
    assert not ops._token_major_ok(x.float(), 1280)
    assert not ops._token_major_ok(torch.zeros(64, 2560, dtype=torch.bfloat16), 1280)  # split storage [hi | lo]
    assert not ops._token_major_ok(torch.zeros(64, 36, dtype=torch.bfloat16)[:, :32], 32)  # row stride not a multiple of 8


def test_the_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import it (a product path routed through the CPU
    restatement would void every parity claim), and the package imports neither triton nor torch.compile."""
    import ast
    from pathlib import Path

    pkg = Path(__file__).resolve().parent.parent / "multimodal-timesfm_b200" / "tsfmx_b200"
    files = sorted(pkg.rglob("*.py"))
    assert len(files) >= 15
    for path in files:
        tree = ast.parse(path.read_text(), str(path))
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            for name in names:
                root = name.split(".")[0]
                assert root not in ("oracle", "triton", "tilelang"), f"{path.name} imports {name}"
            if isinstance(node, ast.Attribute) and node.attr == "compile" and isinstance(node.value, ast.Name):
                assert node.value.id != "torch", f"{path.name} uses torch.compile"
