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
  return 0;
}
"""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_header_is_plain_c_and_links_from_a_c_program(lib, tmp_path):
    """The boundary is a C ABI, not a C++ or torch one: a strict-C99 translation unit that includes the header compiles
    without warnings, links against the in-tree library alone and gets status codes + messages back."""
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
    status, message = run.stdout.strip().split("|", 1)
    assert int(status) == 3 and "no CPU fallback" in message  # TSFMX_ERR_NO_DEVICE
