"""The C-ABI library builds, loads on a CPU-only machine and exports every symbol include/tsfmx_b200.h declares.
No compute call is made here (there is no GPU and no CPU fallback: compute entry points must fail loudly)."""

import re
from pathlib import Path

import pytest
import torch

from tsfmx_b200 import _lib, ops

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    if not _lib.LIB_PATH.exists():
        import __graft_entry__

        __graft_entry__.build()
    return _lib.load()


def declared_symbols():
    text = (ROOT / "include" / "tsfmx_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tsfmx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/tsfmx_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in tsfmx_b200/_lib.py"
    for name in _lib.SIGNATURES:
        assert name in names, f"{name} is bound in _lib.py but not declared in the header"


def test_abi_version_and_error_channel(lib):
    header = (ROOT / "include" / "tsfmx_b200.h").read_text()
    declared = int(re.search(r"#define\s+TSFMX_ABI_VERSION\s+(\d+)", header).group(1))
    assert lib.tsfmx_abi_version() == declared == _lib.ABI_VERSION == 2
    assert isinstance(lib.tsfmx_last_error(), bytes)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    assert lib.tsfmx_device_check(-1) == 3  # TSFMX_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.tsfmx_last_error()
    x = torch.zeros(2, 64)
    m = torch.zeros(2, 64, dtype=torch.bool)
    with pytest.raises(_lib.TsfmxError, match="no CPU fallback"):
        ops.timesfm_patchify_norm(x, m)


def test_gemm_args_struct_layout():
    # must match `tsfmx_gemm_args` in the header (LP64): 2 x 8 + 2 x 40 + ... = 176 bytes
    import ctypes

    assert ctypes.sizeof(_lib.GemmSegment) == 40
    assert ctypes.sizeof(_lib.GemmArgs) == 8 + 4 + 4 + 80 + 4 + 4 + 8 * 4 + 8 + 8 + 8 + 4 * 4 + 8 * 4 + 4 + 4


C_CONSUMER = r"""
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include "tsfmx_b200.h"

int main(void) {
  tsfmx_gemm_args args;
  memset(&args, 0, sizeof(args));
  if (tsfmx_abi_version() != TSFMX_ABI_VERSION) return 10;
  if (tsfmx_sizeof_gemm_args() != (int)sizeof(tsfmx_gemm_args)) return 11;
  if (tsfmx_launch_count() != 0) return 12;
  /* no GPU in this container: a compute entry point must fail with a status and a message, never crash */
  int rc = tsfmx_device_check(-1);
  printf("%d|%s\n", rc, tsfmx_last_error());
#define OFF(T, f) printf(#T "." #f " %d\n", (int)offsetof(T, f))
  printf("tsfmx_gemm_segment.sizeof %d\n", (int)sizeof(tsfmx_gemm_segment));
  OFF(tsfmx_gemm_segment, a); OFF(tsfmx_gemm_segment, lda); OFF(tsfmx_gemm_segment, b); OFF(tsfmx_gemm_segment, ldb);
  OFF(tsfmx_gemm_segment, k);
  printf("tsfmx_gemm_args.sizeof %d\n", (int)sizeof(tsfmx_gemm_args));
  OFF(tsfmx_gemm_args, m); OFF(tsfmx_gemm_args, n); OFF(tsfmx_gemm_args, num_segments); OFF(tsfmx_gemm_args, seg);
  OFF(tsfmx_gemm_args, precision); OFF(tsfmx_gemm_args, act); OFF(tsfmx_gemm_args, bias); OFF(tsfmx_gemm_args, row_scale);
  OFF(tsfmx_gemm_args, row_shift); OFF(tsfmx_gemm_args, residual); OFF(tsfmx_gemm_args, ldr); OFF(tsfmx_gemm_args, d);
  OFF(tsfmx_gemm_args, ldd); OFF(tsfmx_gemm_args, d_dtype); OFF(tsfmx_gemm_args, n_store); OFF(tsfmx_gemm_args, split_off);
  OFF(tsfmx_gemm_args, aux_dtype); OFF(tsfmx_gemm_args, aux); OFF(tsfmx_gemm_args, ld_aux); OFF(tsfmx_gemm_args, pre_act);
  OFF(tsfmx_gemm_args, ld_pre); OFF(tsfmx_gemm_args, pre_act_dtype);
  printf("tsfmx_timesfm_layer.sizeof %d\n", (int)sizeof(tsfmx_timesfm_layer));
  OFF(tsfmx_timesfm_layer, qkv); OFF(tsfmx_timesfm_layer, out); OFF(tsfmx_timesfm_layer, ff0); OFF(tsfmx_timesfm_layer, ff1);
  OFF(tsfmx_timesfm_layer, pre_attn_ln); OFF(tsfmx_timesfm_layer, post_attn_ln); OFF(tsfmx_timesfm_layer, pre_ff_ln);
  OFF(tsfmx_timesfm_layer, post_ff_ln); OFF(tsfmx_timesfm_layer, q_ln); OFF(tsfmx_timesfm_layer, k_ln);
  OFF(tsfmx_timesfm_layer, q_scale);
  printf("tsfmx_timesfm_stack.sizeof %d\n", (int)sizeof(tsfmx_timesfm_stack));
  OFF(tsfmx_timesfm_stack, num_layers); OFF(tsfmx_timesfm_stack, model_dims); OFF(tsfmx_timesfm_stack, num_heads);
  OFF(tsfmx_timesfm_stack, head_dim); OFF(tsfmx_timesfm_stack, ff_dims); OFF(tsfmx_timesfm_stack, precision);
  OFF(tsfmx_timesfm_stack, eps); OFF(tsfmx_timesfm_stack, inv_freq); OFF(tsfmx_timesfm_stack, layers);
  return 0;
}
"""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_header_is_plain_c_and_links_from_a_c_program(lib, tmp_path):
    """The boundary is a C ABI, not a C++ or torch one: a strict-C99 translation unit that includes the header compiles
    without warnings, links against the in-tree library alone and gets status codes + messages back."""
    import ctypes
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "consumer.c"
    src.write_text(C_CONSUMER)
    exe = tmp_path / "consumer"
    r = subprocess.run(
        [gcc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(src), "-o", str(exe),
         f"-L{_lib.LIB_PATH.parent}", "-ltsfmx_b200", f"-Wl,-rpath,{_lib.LIB_PATH.parent}"],
        capture_output=True, text=True,
    )
    assert r.returncode == 0, r.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    first, *layout = run.stdout.strip().splitlines()
    status, message = first.split("|", 1)
    assert int(status) == 3 and "no CPU fallback" in message  # TSFMX_ERR_NO_DEVICE
    # struct layouts as the C compiler sees them == the ctypes Structures of the binding, field by field
    structs = {"tsfmx_gemm_segment": _lib.GemmSegment, "tsfmx_gemm_args": _lib.GemmArgs,
               "tsfmx_timesfm_layer": _lib.TimesfmLayer, "tsfmx_timesfm_stack": _lib.TimesfmStack}
    seen = 0
    for line in layout:
        key, value = line.split()
        cname, field = key.split(".")
        cls = structs[cname]
        want = ctypes.sizeof(cls) if field == "sizeof" else getattr(cls, field).offset
        assert int(value) == want, (key, int(value), want)
        seen += 1
    assert seen == 4 + 5 + 22 + 11 + 9


def _prototypes(text: str, definitions: bool = False):
    """(name -> (return kind, [argument kinds])) of every `tsfmx_*` function declared or defined in C / CUDA source."""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    prefix = r'extern\s+"C"\s+' if definitions else r"(?<![\w\"])"
    for ret, name, args in re.findall(
            prefix + r"((?:const\s+)?[A-Za-z_][A-Za-z0-9_]*\s*\**)\s*\b(tsfmx_[a-z0-9_]+)\s*\(([^;{)]*)\)\s*[;{]", text):
        kinds = []
        args = " ".join(args.split())
        if args not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                ctype = a[: a.rfind(" ")] if " " in a and not a.endswith("*") else a  # drop the parameter name
                if "*" in a:
                    ctype = a[: a.rfind("*") + 1]
                kinds.append(_kind(ctype))
        out[name] = (_kind(ret), kinds)
    return out


def _kind(ctype: str) -> str:
    c = ctype.replace("const", "").replace(" ", "")
    if c.endswith("*"):
        return "char*" if c == "char*" else "ptr"
    return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "float": "f32", "size_t": "u64", "uint64_t": "u64"}[c]  # LP64


def _ctypes_kind(t) -> str:
    import ctypes

    if t in (ctypes.c_void_p,) or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
        return "ptr"
    return {ctypes.c_char_p: "char*", ctypes.c_int32: "i32", ctypes.c_int64: "i64", ctypes.c_float: "f32",
            ctypes.c_size_t: "u64", ctypes.c_uint64: "u64"}[t]


def test_binding_and_definitions_agree_with_the_header_argument_by_argument():
    """Width and order of every argument: header declaration == ctypes signature == the `extern "C"` definition in csrc.
    (A 32- vs 64-bit slip in a binding does not fail at call time; it reads a garbage size.)"""
    header = _prototypes((ROOT / "include" / "tsfmx_b200.h").read_text())
    assert len(header) == len(_lib.SIGNATURES)
    for name, (ret, kinds) in header.items():
        restype, argtypes = _lib.SIGNATURES[name]
        assert _ctypes_kind(restype) == ret, name
        assert [_ctypes_kind(t) for t in argtypes] == kinds, name
    defined = {}
    for src in sorted((ROOT / "multimodal-timesfm_b200" / "csrc").glob("*.cu")):
        for name, proto in _prototypes(src.read_text(), definitions=True).items():
            assert name in header, f"{src.name} defines {name}, which the header does not declare"
            defined[name] = proto
    missing = sorted(set(header) - set(defined))
    assert not missing, missing
    for name, proto in defined.items():
        assert proto == header[name], (name, proto, header[name])


_NO_ARGUMENT_CHECKS = {  # version / accounting queries and test hooks: nothing to refuse
    "tsfmx_abi_version", "tsfmx_last_error", "tsfmx_launch_count", "tsfmx_sizeof_gemm_args", "tsfmx_device_check",
    "tsfmx_gemm_set_cta_group", "tsfmx_gemm_set_split_k", "tsfmx_attention_force_simt", "tsfmx_tune",
}


def test_every_compute_entry_point_refuses_null_arguments_with_a_status(lib):
    """The ABI never throws and never crashes: all-NULL / all-zero arguments come back as TSFMX_ERR_INVALID_ARGUMENT with
    a message, before any device work (SURVEY 8(b) error conventions: errors are raised before any compute)."""
    import ctypes

    def zero(t):
        if t is ctypes.c_float:
            return 0.0
        return None if t in (ctypes.c_void_p, ctypes.c_char_p) or issubclass(t, ctypes._Pointer) else 0

    launches = lib.tsfmx_launch_count()
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        if name in _NO_ARGUMENT_CHECKS:
            continue
        status = getattr(lib, name)(*[zero(t) for t in argtypes])
        if name == "tsfmx_timesfm_stack_workspace_bytes":
            assert status == 0  # size query: 0 bytes for a bad table
        else:
            assert status == 1, (name, status, lib.tsfmx_last_error())  # TSFMX_ERR_INVALID_ARGUMENT
        assert lib.tsfmx_last_error(), name
    assert lib.tsfmx_launch_count() == launches


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_argument_errors_come_before_device_errors(lib):
    """Shape / dtype / alignment errors are reported as such even on a machine without a GPU; well-formed calls fail with
    a device status (never a crash, never a silent success)."""
    import ctypes

    p = 0x10000  # a well-aligned fake device address: never dereferenced on the host
    assert lib.tsfmx_rmsnorm(p, 8, 1000, p, 1e-6, 1, p, None) == 4  # TSFMX_ERR_UNSUPPORTED
    assert b"cols=1000 unsupported" in lib.tsfmx_last_error()
    assert lib.tsfmx_rmsnorm(p + 4, 8, 1280, p, 1e-6, 1, p, None) == 1
    assert b"16-byte aligned" in lib.tsfmx_last_error()
    assert lib.tsfmx_norm_residual_norm(p, 7, p, 8, 1280, p, p, 1e-6, p, 1, p, None) == 1
    # the reference's ValueError for ctx % patch_len != 0 (tsfmx/tsfm/timesfm.py:48-49) exists at the C level too
    assert lib.tsfmx_timesfm_patchify_norm(p, p, 4, 500, 32, 0, p, p, p, p, p, None) == 1
    assert b"must be divisible by patch length" in lib.tsfmx_last_error()
    args = _lib.GemmArgs()
    args.m, args.n, args.num_segments = 128, 128, 1
    args.seg[0].a, args.seg[0].b, args.seg[0].lda, args.seg[0].ldb, args.seg[0].k = p, p, 64, 64, 60
    args.d, args.ldd = p, 128
    assert lib.tsfmx_gemm(ctypes.byref(args), None) == 1
    assert b"multiple of 64" in lib.tsfmx_last_error()
    args.seg[0].k = 64
    assert lib.tsfmx_gemm(ctypes.byref(args), None) in (2, 3)  # TSFMX_ERR_CUDA / TSFMX_ERR_NO_DEVICE
    assert lib.tsfmx_rmsnorm(p, 8, 1280, p, 1e-6, 1, p, None) in (2, 3)
    assert lib.tsfmx_timesfm_patchify_norm(p, p, 4, 512, 32, 0, p, p, p, p, p, None) in (2, 3)


def test_stack_workspace_query_is_pure_host_arithmetic(lib):
    """the caller sizes the scratch of the whole-stack call with a query that needs no device."""
    import ctypes

    layers = (_lib.TimesfmLayer * 50)()
    table = _lib.TimesfmStack(num_layers=50, model_dims=1280, num_heads=16, head_dim=80, ff_dims=1280,
                              precision=_lib.PREC_BF16, eps=1e-6, layers=layers)
    rows = 4096 * 16
    # xn, attn (operands) + qkv, a (GEMM outputs) + h, all bf16 in throughput mode: 7 x rows x 1280 x 2 bytes
    assert lib.tsfmx_timesfm_stack_workspace_bytes(ctypes.byref(table), 4096, 16) == 7 * rows * 1280 * 2
    table.precision = _lib.PREC_BF16X3  # split operands and fp32 intermediates: twice the bytes
    assert lib.tsfmx_timesfm_stack_workspace_bytes(ctypes.byref(table), 4096, 16) == 7 * rows * 1280 * 4
    table.precision = _lib.PREC_BF16
    small = lib.tsfmx_timesfm_stack_workspace_bytes(ctypes.byref(table), 3, 5)  # every buffer rounded up to 256 bytes
    assert small % 256 == 0 and small >= 7 * 15 * 1280 * 2
    assert lib.tsfmx_timesfm_stack_workspace_bytes(ctypes.byref(table), 0, 16) == 0
    table.head_dim = 64  # heads x head_dim != model_dims
    assert lib.tsfmx_timesfm_stack_workspace_bytes(ctypes.byref(table), 4096, 16) == 0
    assert b"heads x head_dim" in lib.tsfmx_last_error()
