"""Multimodal decoder (reference tsfmx/decoder.py:12-92).

Orchestrates ``adapter.preprocess -> fusion -> adapter.forward -> adapter.postprocess``; ``forward``
returns the point-forecast channel.  Same constructor, config dataclass, argument meaning and
``ValueError`` behaviour as the reference; the stages themselves run on the B200 through the C ABI.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import torch
from torch import nn

from . import lanes
from .fusion import MultimodalFusion
from .tsfm.base import TsfmAdapter


@dataclass
class MultimodalDecoderConfig:
    """Configuration for MultimodalDecoder (reference decoder.py:12-18)."""

    text_embedding_dims: int = 384
    num_fusion_layers: int = 1
    fusion_hidden_dims: list[int] = field(default_factory=list)


class MultimodalDecoder(nn.Module):
    """Forecast entry point: time series (+ optional per-patch text embeddings) -> quantile forecasts."""

    def __init__(self, adapter: TsfmAdapter, config: MultimodalDecoderConfig) -> None:
        super().__init__()
        self.adapter = adapter
        self.config = config
        self.fusion = MultimodalFusion(
            ts_embedding_dims=adapter.model_dims,
            text_embedding_dims=config.text_embedding_dims,
            num_layers=config.num_fusion_layers,
            hidden_dims=config.fusion_hidden_dims,
        )
        self.fusion.precision = getattr(adapter, "precision", "bf16")
        # series lanes for forecasting (tsfmx_b200.lanes): 2 = overlap one lane's GEMMs with the other's HBM-bound kernels
        self.lanes = 2
        # opt-in CUDA-graph replay of the forecast path (see _graphed_forecast): the ~360 launches of a 50-layer
        # forward cost ~13 ms of Python per call, which a slow host cannot hide behind an 80 ms step
        self.graphs = False
        self._graph_cache: dict[tuple, tuple] = {}
        self.graph_launches_replayed = 0  # kernels launched through graph replays (each replay = its captured launches)
        self.graph_captures = 0  # graphs captured so far (a caller whose shapes keep changing should stop asking)
        self._lanes_warm_key: tuple | None = None  # (weights, precision, context) the lazily built caches were made for

    def set_precision(self, precision: str) -> None:
        """"bf16" (throughput) or "bf16x3" (parity: <= 1e-3 relative against the fp32 reference)."""
        self.adapter.set_precision(precision)
        self.fusion.precision = precision

    def forward_full(
        self,
        horizon: int,
        inputs: torch.Tensor,
        masks: torch.Tensor,
        text_embeddings: torch.Tensor | None = None,
    ) -> torch.Tensor:
        """(batch, horizon, num_outputs) forecasts; fusion is skipped when ``text_embeddings`` is None.

        Raises ValueError if ``masks.shape != inputs.shape`` (reference decoder.py:62-63).
        """
        if masks.shape != inputs.shape:
            raise ValueError(f"masks shape {masks.shape} must match inputs shape {inputs.shape}")
        masks = masks.bool()
        if inputs.shape[0] == 0:  # empty batch: the reference's eager ops return an empty forecast
            outputs = getattr(self.adapter, "num_outputs", None)
            if outputs is None:
                raise ValueError("empty batch and the adapter does not declare num_outputs")
            return torch.empty(0, horizon, outputs, dtype=torch.float32, device=inputs.device)
        if inputs.is_cuda and inputs.device.index != torch.cuda.current_device():
            # streams, lane streams, graph capture and the library's per-device state all follow the CURRENT device:
            # run the whole call on the device that holds the inputs (a model built with device=cuda:1 in a process
            # that never called torch.cuda.set_device)
            with torch.cuda.device(inputs.device):
                return self._dispatch(horizon, inputs, masks, text_embeddings)
        return self._dispatch(horizon, inputs, masks, text_embeddings)

    def _dispatch(self, horizon, inputs, masks, text_embeddings):
        # train() mode with autograd on = the reference's fine-tune step; eval() / no_grad = plain forecasting
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._forward_full_training(horizon, inputs, masks, text_embeddings)
        if self.graphs and inputs.is_cuda and getattr(self.adapter, "graph_safe", False):
            return self._graphed_forecast(horizon, inputs, masks, text_embeddings)
        return self._forecast(horizon, inputs, masks, text_embeddings)

    GRAPH_CACHE_ENTRIES = 4  # each entry owns the activations of one forward (about 0.7 MB per series at 50 layers)

    def _graphed_forecast(self, horizon, inputs, masks, text_embeddings):
        """Forecast replayed from a CUDA graph.  One graph per (input buffers, shapes, horizon, stream, weights): the
        graph reads the CALLER'S buffers, so a caller that refills the same device tensors batch after batch (the
        evaluator's staging slots) replays one graph per slot.  The returned tensor is the graph's output buffer: it is
        overwritten by the next call with the same key, so consume it (on the same stream) before calling again.
        Parameters are tracked by (data_ptr, version): an optimizer step, ``load_state_dict`` or ``set_precision``
        makes the next call capture afresh."""
        stream = torch.cuda.current_stream(inputs.device)
        key = (
            horizon, inputs.data_ptr(), tuple(inputs.shape), inputs.dtype, masks.data_ptr(),
            0 if text_embeddings is None else text_embeddings.data_ptr(),
            None if text_embeddings is None else (tuple(text_embeddings.shape), text_embeddings.dtype),
            stream.cuda_stream, getattr(self.adapter, "precision", None), self.fusion.precision,
            tuple((p.data_ptr(), p._version) for p in self.parameters()),
        )
        entry = self._graph_cache.get(key)
        if entry is None:
            # eager pass first: argument checks (the reference's ValueErrors), lazily built tables and packed weights
            lanes_before, self.lanes = self.lanes, 1  # single-stream capture; without launch gaps lanes buy nothing
            from . import _lib

            try:
                self._forecast(horizon, inputs, masks, text_embeddings)
                graph = torch.cuda.CUDAGraph()
                launches = _lib.launch_count()
                # thread-local capture mode: CUDA calls of OTHER threads (a DataLoader's pin-memory thread allocating
                # host memory, the bench's clock sampler) must not invalidate the capture; this thread's own calls are
                # checked exactly as in the default mode
                with torch.cuda.graph(graph, stream=torch.cuda.Stream(device=inputs.device),
                                      capture_error_mode="thread_local"):
                    out = self._forecast(horizon, inputs, masks, text_embeddings)
                launches = _lib.launch_count() - launches  # C-ABI kernel launches recorded into the graph
            finally:
                self.lanes = lanes_before
            while len(self._graph_cache) >= self.GRAPH_CACHE_ENTRIES:
                self._graph_cache.pop(next(iter(self._graph_cache)))
            # the inputs are kept alive with the graph: their addresses are baked into it.  So are the packed weight
            # copies the eager pass built: switching the precision mode and back clears the modules' caches, but this
            # entry's key would match again and its replay must still find the weights it was captured with
            packed = [dict(getattr(m, "_packed", {})) for m in (self.adapter, self.fusion)]
            entry = (graph, out, (inputs, masks, text_embeddings), launches, packed)
            self._graph_cache[key] = entry
            self.graph_captures += 1
        entry[0].replay()
        self.graph_launches_replayed += entry[3]
        return entry[1]

    def _caches_key(self, inputs: torch.Tensor, text_embeddings: torch.Tensor | None) -> tuple:
        """What the lazily built device caches depend on: packed bf16 weight copies (adapter and fusion), the Chronos-2
        future-patch embeddings and the RoPE tables are rebuilt when a parameter, the precision, the device or the
        context length changes."""
        return (
            tuple((p.data_ptr(), p._version) for p in self.parameters()),
            getattr(self.adapter, "precision", None), self.fusion.precision, inputs.device, inputs.shape[1],
            text_embeddings is not None,
        )

    def _forecast(self, horizon, inputs, masks, text_embeddings):
        count = self._lane_count(inputs)
        if count > 1:
            # The stages build their caches lazily, i.e. with kernels on whichever stream first asks for them.  In a
            # lane run that is lane 0's stream, and the other lanes would read the half-built tensors from theirs with
            # nothing ordering the two.  So the first forecast after anything the caches depend on has changed runs on
            # the caller's stream alone; the lane streams of every later call wait for that stream on entry.
            key = self._caches_key(inputs, text_embeddings)
            if key != self._lanes_warm_key:
                count = 1
        if count > 1:
            if text_embeddings is not None and text_embeddings.shape[0] != inputs.shape[0]:
                raise ValueError(
                    f"text_embeddings batch ({text_embeddings.shape[0]}) does not match inputs batch ({inputs.shape[0]})"
                )
            cuts = lanes.split_points(inputs.shape[0], count, multiple=8)
            parts = lanes.run_lanes(
                lambda i: self._forecast_steps(
                    horizon,
                    inputs[cuts[i][0] : cuts[i][1]],
                    masks[cuts[i][0] : cuts[i][1]],
                    None if text_embeddings is None else text_embeddings[cuts[i][0] : cuts[i][1]],
                ),
                len(cuts),
                inputs.device,
            )
            return torch.cat(parts, dim=0)
        out = lanes.drain(self._forecast_steps(horizon, inputs, masks, text_embeddings))
        if inputs.is_cuda:
            self._lanes_warm_key = self._caches_key(inputs, text_embeddings)
        return out

    def _forecast_steps(self, horizon, inputs, masks, text_embeddings):
        """preprocess -> fusion -> forward -> postprocess (reference decoder.py:65-72) as one step generator."""
        options = getattr(self.adapter, "forecast_options", None)
        if options is not None and options.active(horizon, getattr(self.adapter._model, "o", horizon)):
            # beyond the reference: upstream's decode loop (horizon > 128) and forecast extras; the text fusion is
            # applied to the context patches (under flip invariance also to those of the negated series)
            fuse = None
            if text_embeddings is not None:
                def fuse(emb):
                    text = text_embeddings
                    if emb.shape[0] == 2 * text.shape[0]:
                        text = torch.cat([text, text], dim=0)
                    return (yield from self.fusion.forward_device_steps(emb, text))
            return (yield from self.adapter.decode_steps(horizon, inputs, masks, fuse))
        preprocessed = yield from lanes.steps(self.adapter, "preprocess", inputs, masks)
        if text_embeddings is not None:
            if hasattr(self.fusion, "forward_device_steps") and not self.fusion.training:
                embeddings = yield from self.fusion.forward_device_steps(preprocessed.input_embeddings, text_embeddings)
            else:
                embeddings = self.fusion(preprocessed.input_embeddings, text_embeddings)
                yield
        else:
            embeddings = preprocessed.input_embeddings
        output_embeddings = yield from lanes.steps(self.adapter, "forward", embeddings, preprocessed.masks)
        return (
            yield from lanes.steps(self.adapter, "postprocess", horizon, output_embeddings, preprocessed.normalization_stats)
        )

    def _lane_count(self, inputs: torch.Tensor) -> int:
        """Lanes to cut a forecast batch into: ``self.lanes`` when every lane still fills the GPU (>= 8192 patch
        tokens, i.e. >= 64 GEMM row tiles of 128), else 1."""
        want = int(getattr(self, "lanes", 1))
        if want <= 1 or not inputs.is_cuda or inputs.dim() != 2:
            return 1
        tokens = inputs.shape[0] * max(1, inputs.shape[1] // max(1, self.adapter.patch_len))
        return max(1, min(want, tokens // 8192))

    def _forward_full_training(self, horizon, inputs, masks, text_embeddings):
        """Differentiable paths of the reference's two training modes: "multimodal" (trainer.py:76-77,119-123: frozen
        adapter, trainable fusion -> ``FusedForecastFunction``) and "baseline" (trainer.py:78-79: the whole adapter is
        trained -> ``FullFineTuneFunction``, for adapters that provide ``preprocess_backward``)."""
        from .autograd import FullFineTuneFunction, FusedForecastFunction

        if any(p.requires_grad for p in self.adapter.parameters()):
            # "baseline" mode of the reference (trainer.py:78-79,123): the whole adapter is trained
            if not hasattr(self.adapter, "preprocess_backward"):
                raise NotImplementedError(
                    f"{type(self.adapter).__name__} has no full fine-tuning path; freeze it to train the fusion module"
                )
            named = [(k, v) for k, v in self.adapter._model.named_parameters()]
            fusion_w = [lin.weight for lin in self.fusion.linears()] if text_embeddings is not None else []
            return FullFineTuneFunction.apply(self, horizon, inputs, masks, text_embeddings, tuple(k for k, _ in named),
                                              len(fusion_w), *fusion_w, *[v for _, v in named])
        if text_embeddings is None:
            raise ValueError("training the fusion module needs text_embeddings")
        if not hasattr(self.adapter, "forward_saving"):
            raise NotImplementedError(f"{type(self.adapter).__name__} has no training path yet")
        weights = [lin.weight for lin in self.fusion.linears()]
        return FusedForecastFunction.apply(self, horizon, inputs, masks, text_embeddings, *weights)

    def forward(
        self,
        horizon: int,
        inputs: torch.Tensor,
        masks: torch.Tensor,
        text_embeddings: torch.Tensor | None = None,
    ) -> torch.Tensor:
        """(batch, horizon) point forecast = ``forward_full(...)[..., adapter.point_forecast_index]``."""
        return self.forward_full(horizon, inputs, masks, text_embeddings)[..., self.adapter.point_forecast_index]
