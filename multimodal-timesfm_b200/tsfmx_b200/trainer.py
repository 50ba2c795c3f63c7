"""Fine-tune entry point (reference tsfmx/trainer.py:35-399): MSE on the point forecast, gradient accumulation,
gradient clipping, AdamW, per-step LR schedule.

Kept from the reference: the class name, constructor arguments, "multimodal" = frozen adapter + trainable fusion
(trainer.py:76-77,119-123), loss / accumulation / clip / step order (trainer.py:200-219), ``RuntimeError`` on empty
datasets, checkpoint dictionary keys (types.py:42-61), the warm-up semantics of ``TrainingArguments.warmup_steps``
(training_args.py:111-121).  Added for the B200 box: data parallelism — each rank takes its slice of every batch,
weights its loss by slice size / batch size and the gradients are summed with NCCL all-reduces per optimizer step,
before the clip (equal to the single-process global-batch step for any slice sizes, empty ones included); the slice of
the next batch is copied host -> device on a copy stream while the current one computes; checkpoints are written by
rank 0.  "baseline" = full fine-tuning of the adapter without text (trainer.py:78-79): available for both reference
adapters, TimesFM 2.5 and Chronos-2 (weight-gradient GEMMs for every Linear, every gradient all-reduced, layer by layer
under the backward pass).
"""

from __future__ import annotations

import math
from collections.abc import Iterator, Sequence
from typing import Any

import torch
from torch import nn
from torch.optim import AdamW, Optimizer
from torch.optim.lr_scheduler import LambdaLR, LRScheduler
from torch.utils.data import DataLoader

from . import distributed as tdist
from .data.collate import baseline_collate_fn, multimodal_collate_fn
from .decoder import MultimodalDecoder


def linear_schedule_with_warmup(optimizer: Optimizer, warmup_steps: int, total_steps: int) -> LambdaLR:
    """Linear warm-up then linear decay to zero (reference optimization.py:19-45)."""

    def fn(step: int) -> float:
        if step < warmup_steps:
            return step / max(1, warmup_steps)
        return max(0.0, (total_steps - step) / max(1, total_steps - warmup_steps))

    return LambdaLR(optimizer, fn)


def cosine_schedule_with_warmup(optimizer: Optimizer, warmup_steps: int, total_steps: int, cycles: float = 0.5) -> LambdaLR:
    """Linear warm-up then cosine decay (reference optimization.py:48-79)."""

    def fn(step: int) -> float:
        if step < warmup_steps:
            return step / max(1, warmup_steps)
        progress = (step - warmup_steps) / max(1, total_steps - warmup_steps)
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * cycles * 2.0 * progress)))

    return LambdaLR(optimizer, fn)


class MultimodalTrainer:
    """Trainer for the fusion fine-tune ("multimodal" mode) and the full fine-tune ("baseline") on one or several B200s."""

    def __init__(
        self,
        model: MultimodalDecoder,
        args: Any,
        train_dataset: Sequence,
        val_dataset: Sequence,
        mode: str,
        device: torch.device,
        wandb_run: Any | None = None,
        optimizers: tuple[Optimizer | None, LRScheduler | None] = (None, None),
    ) -> None:
        self.model = model
        self.args = args
        self.train_dataset = train_dataset
        self.val_dataset = val_dataset
        self.mode = mode
        self.device = device
        self._wandb_run = wandb_run
        self.rank, self.world_size, _ = tdist.world_info()
        self.model.to(self.device)
        if mode == "multimodal":
            self.model.adapter.freeze_parameters()
        elif mode == "baseline":
            # full fine-tuning of the adapter, no text (reference trainer.py:78-79, baseline_collate_fn)
            # the C-ABI adapters compute their gradients by hand: one without the weight-gradient path would silently
            # train nothing (a plain autograd adapter, e.g. the CPU oracle, needs no such path)
            from .tsfm.base import TsfmAdapter

            if isinstance(self.model.adapter, TsfmAdapter) and not hasattr(self.model.adapter, "preprocess_backward"):
                raise NotImplementedError(
                    f"mode='baseline' needs backbone weight gradients, which {type(self.model.adapter).__name__} does "
                    "not produce on the B200 path (TimesFM2p5Adapter and Chronos2Adapter do)"
                )
            self.model.adapter.unfreeze_parameters()
        else:
            raise ValueError(f"mode must be 'multimodal' or 'baseline', got {mode!r}")
        collate = multimodal_collate_fn if mode == "multimodal" else baseline_collate_fn
        pin = self.device.type == "cuda"
        self.train_loader = DataLoader(train_dataset, batch_size=args.per_device_train_batch_size * self.world_size,
                                       shuffle=True, num_workers=0, collate_fn=collate, pin_memory=pin,
                                       generator=torch.Generator().manual_seed(getattr(args, "seed", 0)))
        self.val_loader = DataLoader(val_dataset, batch_size=args.per_device_eval_batch_size * self.world_size,
                                     shuffle=False, num_workers=0, collate_fn=collate, pin_memory=pin)
        self.loss_fn = nn.MSELoss()
        optimizer, scheduler = optimizers
        steps_per_epoch = math.ceil(len(self.train_loader) / args.gradient_accumulation_steps)
        total_steps = args.num_train_epochs * steps_per_epoch
        self.optimizer = optimizer or AdamW(self._get_trainable_params(), lr=args.learning_rate,
                                            weight_decay=args.weight_decay)
        self.lr_scheduler = scheduler or self._create_scheduler(total_steps)
        self.current_epoch = 0
        self.global_step = 0
        self.best_val_loss = float("inf")
        self._copy_stream: torch.cuda.Stream | None = None
        self._micro_batches_since_step = 0
        # CUDA-graph replay of forward + loss + backward for micro-batch shapes that repeat (every full batch of an
        # epoch): a full fine-tune step is ~5 800 kernel launches and was bound by the host's launch rate.  The optimizer
        # step (all-reduce, clip, AdamW) stays outside the graph.  Not with gradient accumulation (a replay overwrites
        # the gradients of the previous micro-batch) and not around the overlapped NCCL reducer.
        self.graphs = bool(getattr(args, "cuda_graphs", self.device.type == "cuda")) and self.device.type == "cuda"
        self._train_graphs: dict[tuple, "_GraphedMicroBatch"] = {}
        self._graph_seen: dict[tuple, int] = {}
        self.graph_replays = 0
        # full fine-tuning on several GPUs: all-reduce each layer's gradients while the backward pass continues below
        # it (1.99 GB per step at 50 layers).  Only without gradient accumulation - with it the collective runs once
        # per optimizer step on the accumulated gradients instead of once per micro-batch.
        self.model.grad_ready_hook = None
        if (mode == "baseline" and self.world_size > 1 and args.gradient_accumulation_steps == 1
                and getattr(args, "overlap_grad_allreduce", True)):
            self.model.grad_ready_hook = tdist.OverlappedGradReducer()

    def _get_trainable_params(self) -> Iterator[nn.Parameter]:
        """Fusion weights in "multimodal" mode, the adapter's trainable parameters in "baseline" mode
        (reference trainer.py:119-123)."""
        if self.mode == "multimodal":
            return self.model.fusion.parameters()
        return (p for p in self.model.adapter.parameters() if p.requires_grad)

    def _warmup_steps(self, total_steps: int) -> int:
        """The reference's ``TrainingArguments.warmup_steps`` is a float: values >= 1 are absolute optimizer steps,
        values in [0, 1) a ratio of the total (``ceil(total * ratio)``, reference training_args.py:111-121)."""
        getter = getattr(self.args, "get_warmup_steps", None)
        if callable(getter):
            return int(getter(total_steps))
        ws = float(getattr(self.args, "warmup_steps", 0.0) or 0.0)
        return int(ws) if ws >= 1 else math.ceil(total_steps * ws)

    def _create_scheduler(self, total_steps: int) -> LRScheduler:
        kind = getattr(self.args, "lr_scheduler_type", "linear")
        warmup = self._warmup_steps(total_steps)
        if kind == "linear":
            return linear_schedule_with_warmup(self.optimizer, warmup, total_steps)
        if kind == "cosine":
            return cosine_schedule_with_warmup(self.optimizer, warmup, total_steps)
        raise NotImplementedError(f"Unknown lr_scheduler_type: {kind}")

    # ------------------------------------------------------------------ data path: shard, stage one batch ahead
    _KEYS = ("context", "horizon", "text_embeddings")

    def _staged(self, loader) -> Iterator[dict]:
        """Yield this rank's shard of every global batch, already on the device.

        On a CUDA device the shard of batch i + 1 is copied (pinned host memory -> HBM, ``non_blocking``) on a copy
        stream while the kernels of batch i run: the reference's loop (trainer.py:200-206) does three synchronous
        ``.to(device)`` per micro-batch in front of the forward.  Each yielded dict also carries ``global_size``, the
        number of samples of the global batch the shard was cut from."""
        cuda = self.device.type == "cuda"
        if not cuda:
            for batch in loader:
                n = len(batch["context"])
                shard = tdist.shard_batch(batch, self.rank, self.world_size)
                out = {k: shard[k].to(self.device) for k in self._KEYS if k in shard}
                out["global_size"] = n
                yield out
            return
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        copy = self._copy_stream

        def stage(batch: dict):
            n = len(batch["context"])
            shard = tdist.shard_batch(batch, self.rank, self.world_size)
            out: dict[str, Any] = {"global_size": n}
            with torch.cuda.stream(copy):
                for k in self._KEYS:
                    if k in shard:
                        dst = shard[k].to(self.device, non_blocking=True)  # allocated on the copy stream ...
                        dst.record_stream(main)  # ... consumed on the caller's: keep the block until main is done
                        out[k] = dst
                done = torch.cuda.Event()
                done.record(copy)
            return out, done

        it = iter(loader)
        try:
            pending = stage(next(it))
        except StopIteration:
            return
        while pending is not None:
            cur, done = pending
            try:
                pending = stage(next(it))
            except StopIteration:
                pending = None
            main.wait_event(done)
            yield cur

    # ------------------------------------------------------------------ one micro-batch / one optimizer step
    def _forward_loss(self, batch: dict) -> torch.Tensor:
        """This rank's share of the global-batch MSE: ``mse(shard) * shard_size / global_size``.  Summed over ranks
        that is exactly the reference's ``MSELoss`` over the whole batch (trainer.py:105,208), whatever the shard sizes
        are - a ragged last batch may leave ranks with fewer samples than the others, or with none (their share is a
        constant zero)."""
        context = batch["context"]
        n_local, n_global = context.shape[0], batch["global_size"]
        if n_local == 0:
            return torch.zeros((), dtype=torch.float32, device=self.device)
        horizon = batch["horizon"]
        input_padding = torch.zeros_like(context, dtype=torch.bool)  # reference trainer.py:204
        point = self.model(horizon.shape[-1], context, input_padding, batch.get("text_embeddings"))
        loss = self.loss_fn(point, horizon)
        return loss if n_local == n_global else loss * (n_local / n_global)

    MAX_TRAIN_GRAPHS = 2  # full batches (and at most one other repeating shape); each graph owns a step's activations

    def _graph_key(self, batch: dict) -> tuple | None:
        if not self.graphs or self.args.gradient_accumulation_steps != 1:
            return None
        if self.model.grad_ready_hook is not None and not getattr(self.args, "graph_collectives", False):
            return None  # the overlapped NCCL all-reduces would be captured too: opt-in (args.graph_collectives)
        if batch["context"].shape[0] == 0 or not getattr(self.model.adapter, "graph_safe", False):
            return None
        return tuple((k, tuple(batch[k].shape)) for k in self._KEYS if k in batch) + (batch["global_size"],)

    def _micro_batch(self, batch: dict, accum: int) -> torch.Tensor:
        """Forward, loss and backward of one micro-batch -> the (detached) loss.  The second time a micro-batch shape
        shows up its forward + backward are captured into a CUDA graph and replayed from then on."""
        key = self._graph_key(batch)
        if key is not None:
            entry = self._train_graphs.get(key)
            if entry is None and len(self._train_graphs) < self.MAX_TRAIN_GRAPHS:
                self._graph_seen[key] = self._graph_seen.get(key, 0) + 1
                if self._graph_seen[key] >= 2:  # the first occurrence ran eagerly and built every lazy cache
                    entry = self._train_graphs[key] = _GraphedMicroBatch(self, batch)
            if entry is not None:
                self._micro_batches_since_step += 1
                self.graph_replays += 1
                return entry.run(batch)
        loss = self._forward_loss(batch) / accum
        self._backward(loss, batch["global_size"])
        return loss.detach() * accum

    def release_graphs(self) -> None:
        """Drop the captured training graphs (and the activations they pin).  With ``args.graph_collectives`` the graphs
        hold captured NCCL kernels: call this BEFORE ``torch.distributed.destroy_process_group()`` - tearing the
        communicator down while such a graph is alive hangs the process."""
        self._train_graphs.clear()
        self._graph_seen.clear()
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def _backward(self, loss: torch.Tensor, global_size: int | None = None) -> None:
        """``loss.backward()``.  A global batch with fewer samples than ranks leaves some rank with an empty shard, a
        constant loss and no backward pass; every rank can tell from ``global_size`` alone, so for such a batch ALL
        ranks skip the overlapped collectives and the gradients are summed in ``optimizer_step`` instead."""
        self._micro_batches_since_step += 1
        reducer = getattr(self.model, "grad_ready_hook", None)
        overlap = reducer is not None and (global_size is None or global_size >= self.world_size)
        if reducer is not None and not overlap:
            self.model.grad_ready_hook = None
        try:
            if loss.requires_grad:
                loss.backward()
        finally:
            if reducer is not None and not overlap:
                self.model.grad_ready_hook = reducer
                reducer.reduced = False

    def optimizer_step(self) -> None:
        """All-reduce (sum of the ranks' shares) of the gradients, clip, AdamW, LR schedule (reference
        trainer.py:213-219).  The all-reduce comes before the clip so that the clip sees the global-batch norm."""
        params = list(self._get_trainable_params())
        reducer = getattr(self.model, "grad_ready_hook", None)
        if self.world_size > 1 and reducer is not None and reducer.reduced and self._micro_batches_since_step == 1:
            reducer.reduced = False  # the backward pass already summed the gradients over ranks, layer by layer
        elif self.world_size > 1:
            for p in params:  # a rank whose shards were all empty still takes part in the collective
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            tdist.allreduce_([p.grad for p in params], "sum")
        self._micro_batches_since_step = 0
        params = [p for p in params if p.grad is not None]
        if self.args.max_grad_norm > 0:
            nn.utils.clip_grad_norm_(params, self.args.max_grad_norm)
        self.optimizer.step()
        self.optimizer.zero_grad()
        self.lr_scheduler.step()
        self.global_step += 1

    def train_epoch(self) -> float:
        """Average training loss of the epoch (global-batch MSE per micro-batch, averaged over micro-batches).

        Raises RuntimeError if the training dataset is empty (reference trainer.py:196-197)."""
        self.model.train()
        num_batches = len(self.train_loader)
        if num_batches == 0:
            raise RuntimeError("Training dataset is empty.")
        accum = self.args.gradient_accumulation_steps
        losses = []
        for i, batch in enumerate(self._staged(self.train_loader)):
            losses.append(self._micro_batch(batch, accum))  # no per-micro-batch .item() sync (reference trainer.py:211)
            if (i + 1) % accum == 0 or (i + 1) == num_batches:
                self.optimizer_step()
        total = torch.stack(losses).sum()
        if self.world_size > 1:
            tdist.allreduce_([total], "sum")
        return float(total.item()) / num_batches

    def validate_epoch(self) -> float:
        """Raises RuntimeError if the validation dataset is empty (reference trainer.py:258-259)."""
        self.model.eval()
        num_batches = len(self.val_loader)
        if num_batches == 0:
            raise RuntimeError("Validation dataset is empty.")
        with torch.no_grad():
            total = torch.stack([self._forward_loss(batch) for batch in self._staged(self.val_loader)]).sum()
        if self.world_size > 1:
            tdist.allreduce_([total], "sum")
        return float(total.item()) / num_batches

    def train(self) -> None:
        """Epoch loop of the reference (trainer.py:356-399): train, validate, log, checkpoint, optionally reload the best
        fusion weights at the end.  Raises NotImplementedError for eval strategies other than "epoch"."""
        strategy = getattr(self.args, "eval_strategy", "epoch")
        if strategy != "epoch":
            raise NotImplementedError(f"eval_strategy={strategy!r} is not supported; only 'epoch' is implemented.")
        for epoch in range(self.args.num_train_epochs):
            self.current_epoch = epoch
            epoch_lr = self.optimizer.param_groups[0]["lr"]
            train_loss = self.train_epoch()
            val_loss = self.validate_epoch()
            if self._wandb_run is not None and self.rank == 0:
                if getattr(self.args, "logging_strategy", "epoch") == "epoch":
                    self._wandb_run.log({"train/loss": train_loss, "train/lr": epoch_lr, "val/loss": val_loss},
                                        step=self.global_step)
                else:
                    self._wandb_run.log({"val/loss": val_loss}, step=self.global_step)
            if getattr(self.args, "save_strategy", "no") in ("epoch", "best"):
                self.save_checkpoint(val_loss)
            else:
                self.best_val_loss = min(self.best_val_loss, val_loss)
        if getattr(self.args, "load_best_model_at_end", False):
            best = self.args.checkpoint_dir / "best_model.pt"
            world_barrier()  # rank 0 has finished writing
            if best.exists():
                self._load_checkpoint_state(torch.load(best, weights_only=True, map_location=self.device))

    # ------------------------------------------------------------------ checkpoints (rank 0 writes; every rank reads)
    def build_checkpoint(self) -> dict[str, Any]:
        """Same keys as the reference ``MultimodalCheckpoint`` / ``BaselineCheckpoint`` (types.py:42-61, trainer.py:285-303)."""
        ckpt = {
            "epoch": self.current_epoch,
            "global_step": self.global_step,
            "optimizer_state_dict": self.optimizer.state_dict(),
            "scheduler_state_dict": self.lr_scheduler.state_dict(),
            "best_val_loss": self.best_val_loss,
        }
        if self.mode == "multimodal":
            ckpt["fusion_state_dict"] = self.model.fusion.state_dict()
        else:
            ckpt["adapter_state_dict"] = self.model.adapter.state_dict()
        return ckpt

    _build_checkpoint = build_checkpoint  # the reference's (private) name

    def _load_checkpoint_state(self, checkpoint: dict[str, Any]) -> None:
        """Restore the trained module from a checkpoint dict (reference trainer.py:305-310)."""
        if self.mode == "multimodal":
            self.model.fusion.load_state_dict(checkpoint["fusion_state_dict"])
        else:
            self.model.adapter.load_state_dict(checkpoint["adapter_state_dict"])

    def _rotate_checkpoints(self) -> None:
        """Keep the ``save_total_limit`` most recent ``checkpoint_epoch_*.pt`` files (reference trainer.py:312-323)."""
        limit = getattr(self.args, "save_total_limit", None)
        if limit is None:
            return
        files = sorted(self.args.checkpoint_dir.glob("checkpoint_epoch_*.pt"), key=lambda f: int(f.stem.rsplit("_", 1)[-1]))
        for stale in files[: max(0, len(files) - limit)]:
            stale.unlink()

    def save_checkpoint(self, val_loss: float) -> None:
        """"epoch": write ``checkpoint_epoch_{n}.pt`` every epoch (rotated) and ``best_model.pt`` on improvement;
        "best": write ``best_model.pt`` only on improvement (reference trainer.py:325-354).  Data-parallel runs hold
        identical weights on every rank, so only rank 0 touches the file system."""
        is_best = val_loss < self.best_val_loss
        if is_best:
            self.best_val_loss = val_loss
        strategy = getattr(self.args, "save_strategy", "epoch")
        if strategy == "best" and not is_best:
            return
        if self.rank != 0:
            return
        checkpoint = self.build_checkpoint()
        self.args.checkpoint_dir.mkdir(parents=True, exist_ok=True)
        if strategy == "epoch":
            torch.save(checkpoint, self.args.checkpoint_dir / f"checkpoint_epoch_{self.current_epoch}.pt")
            self._rotate_checkpoints()
        if is_best:
            torch.save(checkpoint, self.args.checkpoint_dir / "best_model.pt")


class _GraphedMicroBatch:
    """Forward + loss + backward of one micro-batch shape as a CUDA graph over static input buffers.

    Captured with the gradients unset, so the backward pass allocates every ``.grad`` from the graph's private pool;
    a replay overwrites them in place, and ``run`` re-attaches them to the parameters (``zero_grad`` drops them after
    every optimizer step).  Packed bf16 copies of weights that changed since the last step are rebuilt by kernels
    inside the graph - they read the parameters at their (fixed) addresses."""

    def __init__(self, trainer: "MultimodalTrainer", batch: dict) -> None:
        self.static = {k: batch[k].clone() for k in trainer._KEYS if k in batch}
        self.static["global_size"] = batch["global_size"]
        params = list(trainer._get_trainable_params())
        trainer.optimizer.zero_grad(set_to_none=True)
        torch.cuda.synchronize(trainer.device)
        self.graph = torch.cuda.CUDAGraph()
        # thread-local error mode: the DataLoader's pin-memory thread may allocate page-locked memory meanwhile; the
        # backward pass runs on autograd's worker thread but on the capturing stream, so it is recorded all the same
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            loss = trainer._forward_loss(self.static)
            loss.backward()
        self.loss = loss.detach()
        # the graph reads packed weight copies that were built BEFORE the capture (a frozen adapter's, cached by the eager
        # first step) at their addresses: hold them, so that clearing a cache later cannot free memory a replay still reads
        self.keepalive = [dict(getattr(m, "_packed", {})) for m in (trainer.model.adapter, trainer.model.fusion)]
        # with the overlapped reducer active the per-layer all-reduces were captured as well: a replay leaves globally
        # summed gradients behind, which optimizer_step has to be told
        self.reducer = trainer.model.grad_ready_hook
        self.grads = [(p, p.grad) for p in params if p.grad is not None]
        for p, _ in self.grads:  # nothing has run yet: the first replay produces the values
            p.grad = None

    def run(self, batch: dict) -> torch.Tensor:
        for k, buf in self.static.items():
            if k != "global_size":
                buf.copy_(batch[k], non_blocking=True)
        self.graph.replay()
        for p, g in self.grads:
            p.grad = g
        if self.reducer is not None:
            self.reducer.reduced = True
        return self.loss.clone()


def world_barrier() -> None:
    """Barrier across ranks when a process group exists (checkpoint files are written by rank 0 only)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()
