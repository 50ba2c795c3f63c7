"""ctypes binding of the tsfmx_b200 C ABI (``include/tsfmx_b200.h``).

The shared library is built in-tree by ``csrc/build.py`` (nvcc, sm_100a).  There is no CPU fallback:
if the library is missing, or a compute entry point is called without a B200, this module raises.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint64, c_void_p
from pathlib import Path

import torch

_CSRC = Path(__file__).resolve().parent.parent / "csrc"
LIB_PATH = _CSRC / "libtsfmx_b200.so"

OK = 0
ABI_VERSION = 2  # TSFMX_ABI_VERSION of include/tsfmx_b200.h this binding was written against
PREC_BF16, PREC_BF16X3 = 0, 1
DT_F32, DT_BF16, DT_BF16_SPLIT = 0, 1, 2
ACT_NONE, ACT_SILU, ACT_RELU, ACT_SILU_GRAD, ACT_RELU_GRAD = 0, 1, 2, 3, 4

MAX_KV_REGIONS = 16  # TSFMX_MAX_KV_REGIONS

PRECISIONS = {"bf16": PREC_BF16, "bf16x3": PREC_BF16X3}


class TsfmxError(RuntimeError):
    """A tsfmx_b200 entry point returned a non-zero status."""


class GemmSegment(Structure):
    _fields_ = [
        ("a", c_void_p),
        ("lda", c_int64),
        ("b", c_void_p),
        ("ldb", c_int64),
        ("k", c_int32),
        ("reserved", c_int32),
    ]


class GemmArgs(Structure):
    _fields_ = [
        ("m", c_int64),
        ("n", c_int32),
        ("num_segments", c_int32),
        ("seg", GemmSegment * 2),
        ("precision", c_int32),
        ("act", c_int32),
        ("bias", c_void_p),
        ("row_scale", c_void_p),
        ("row_shift", c_void_p),
        ("residual", c_void_p),
        ("ldr", c_int64),
        ("d", c_void_p),
        ("ldd", c_int64),
        ("d_dtype", c_int32),
        ("n_store", c_int32),
        ("split_off", c_int32),
        ("aux_dtype", c_int32),
        ("aux", c_void_p),
        ("ld_aux", c_int64),
        ("pre_act", c_void_p),
        ("ld_pre", c_int64),
        ("pre_act_dtype", c_int32),
        ("reserved", c_int32),
    ]


class TimesfmLayer(Structure):
    _fields_ = [(name, c_void_p) for name in ("qkv", "out", "ff0", "ff1", "pre_attn_ln", "post_attn_ln", "pre_ff_ln",
                                              "post_ff_ln", "q_ln", "k_ln", "q_scale")]


class TimesfmStack(Structure):
    _fields_ = [
        ("num_layers", c_int32), ("model_dims", c_int32), ("num_heads", c_int32), ("head_dim", c_int32),
        ("ff_dims", c_int32), ("precision", c_int32), ("eps", c_float), ("reserved", c_int32),
        ("inv_freq", c_void_p), ("layers", POINTER(TimesfmLayer)),
    ]


# name -> (restype, argtypes); every symbol declared in include/tsfmx_b200.h
SIGNATURES: dict[str, tuple[object, list[object]]] = {
    "tsfmx_abi_version": (c_int32, []),
    "tsfmx_last_error": (c_char_p, []),
    "tsfmx_launch_count": (c_uint64, []),
    "tsfmx_device_check": (c_int32, [c_int32]),
    "tsfmx_timesfm_patchify_norm": (
        c_int32,
        [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
         c_void_p],
    ),
    "tsfmx_chronos2_patchify_norm": (
        c_int32,
        [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_float, c_int32, c_int32, c_void_p, c_void_p,
         c_void_p, c_void_p, c_void_p],
    ),
    "tsfmx_chronos_t5_tokenize": (
        c_int32,
        [c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
         c_void_p, c_void_p],
    ),
    "tsfmx_chronos_t5_dequantize": (
        c_int32,
        [c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p],
    ),
    "tsfmx_cast_rows": (c_int32, [c_void_p, c_int64, c_int32, c_int64, c_int32, c_void_p, c_void_p]),
    "tsfmx_gemm": (c_int32, [POINTER(GemmArgs), c_void_p]),
    "tsfmx_gemm_set_cta_group": (c_int32, [c_int32]),
    "tsfmx_gemm_set_split_k": (c_int32, [c_int32]),
    "tsfmx_gemm_wgrad": (
        c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p, c_int64, c_void_p],
    ),
    "tsfmx_gemm_rownorm": (
        c_int32,
        [POINTER(GemmSegment), c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p,
         c_float, c_void_p],
    ),
    "tsfmx_rmsnorm": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_float, c_int32, c_void_p, c_void_p]),
    "tsfmx_norm_residual_norm": (
        c_int32,
        [c_void_p, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_float, c_void_p, c_int32, c_void_p,
         c_void_p],
    ),
    "tsfmx_attention_force_simt": (c_int32, [c_int32]),
    "tsfmx_tune": (c_int32, [c_int32, c_int32]),
    "tsfmx_t5_sample_topk": (
        c_int32, [c_void_p, c_int64, c_int32, c_int32, c_float, c_int32, c_void_p, c_void_p, c_void_p]
    ),
    "tsfmx_embed_rows": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "tsfmx_t5_attention": (
        c_int32,
        [c_void_p, c_int32, c_int64, c_int64, c_void_p, c_void_p, c_int32, c_int64, c_int64, c_int64, c_int64, c_int32,
         c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int64,
         c_int64, c_int32, c_void_p],
    ),
    "tsfmx_t5_encoder_attention_mma": (
        c_int32,
        [c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "tsfmx_colsum_wgrad": (
        c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_int64, c_int32, c_float, c_int32, c_void_p, c_void_p]
    ),
    "tsfmx_encoder_attention_bwd": (
        c_int32,
        [c_void_p, c_int32, c_void_p, c_int32, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32,
         c_void_p, c_void_p],
    ),
    "tsfmx_rope_table": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "tsfmx_encoder_attention_mma": (
        c_int32,
        [c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "tsfmx_encoder_attention": (
        c_int32,
        [c_void_p, c_int32, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_void_p],
    ),
    "tsfmx_chronos2_finalize": (
        c_int32,
        [c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "tsfmx_sizeof_gemm_args": (c_int32, []),
    "tsfmx_rmsnorm_bwd_chain": (
        c_int32,
        [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int64, c_int32,
         c_float, c_void_p, c_int32, c_void_p, c_void_p],
    ),
    "tsfmx_rmsnorm_bwd_chain_wgrad": (
        c_int32,
        [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int64, c_int32,
         c_float, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "tsfmx_timesfm_attention_bwd": (
        c_int32,
        [c_void_p, c_int32, c_void_p, c_int32, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
         c_void_p, c_void_p, c_void_p, c_float, c_int32, c_void_p, c_void_p, c_void_p],
    ),
    "tsfmx_transpose_mask": (
        c_int32,
        [c_void_p, c_int32, c_int64, c_int32, c_int64, c_void_p, c_int32, c_int64, c_int32, c_void_p, c_int64,
         c_void_p],
    ),
    "tsfmx_mask_cast_rows": (
        c_int32,
        [c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int64, c_int32, c_void_p, c_void_p],
    ),
    "tsfmx_timesfm_stack_workspace_bytes": (c_size_t, [POINTER(TimesfmStack), c_int64, c_int32]),
    "tsfmx_timesfm_stack_fwd": (
        c_int32,
        [POINTER(TimesfmStack), c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p],
    ),
    "tsfmx_timesfm_patchify_continue": (
        c_int32,
        [c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_void_p,
         c_void_p, c_void_p, c_void_p],
    ),
    "tsfmx_timesfm_attention_decode": (
        c_int32,
        [POINTER(c_void_p), POINTER(c_int32), c_int32, c_int32, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p,
         c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int32, c_void_p, c_void_p],
    ),
    "tsfmx_timesfm_forecast_finalize": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
         c_int32, c_void_p, c_void_p],
    ),
    "tsfmx_timesfm_attention": (
        c_int32,
        [c_void_p, c_int32, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
         c_void_p, c_float, c_int32, c_void_p, c_void_p],
    ),
}

_lib: ctypes.CDLL | None = None


def load() -> ctypes.CDLL:
    """Load the in-tree shared library (loudly fails if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("TSFMX_B200_LIB", LIB_PATH))
    if not path.exists():
        raise TsfmxError(
            f"{path} not found: build the CUDA extension first (python __graft_entry__.py build, or "
            f"python {_CSRC / 'build.py'}). tsfmx_b200 has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(str(path))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.tsfmx_abi_version() != ABI_VERSION:
        raise TsfmxError(f"ABI version mismatch: library reports {lib.tsfmx_abi_version()}, binding expects {ABI_VERSION}")
    if lib.tsfmx_sizeof_gemm_args() != ctypes.sizeof(GemmArgs):
        raise TsfmxError(
            f"tsfmx_gemm_args layout mismatch: library {lib.tsfmx_sizeof_gemm_args()} bytes, binding "
            f"{ctypes.sizeof(GemmArgs)} bytes"
        )
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != OK:
        msg = load().tsfmx_last_error()
        raise TsfmxError(f"tsfmx_b200 call failed (status {status}): {msg.decode() if msg else '?'}")


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors: torch.Tensor | None) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise TsfmxError(
                "tsfmx_b200 runs on B200 only: got a tensor on " f"{t.device}; there is no CPU fallback"
            )


def launch_count() -> int:
    return int(load().tsfmx_launch_count())
