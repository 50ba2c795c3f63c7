"""Build libtsfmx_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

Usage: python build.py [--force] [--verbose]
The shared library is written next to the sources so that it travels with the repo snapshot.
"""

from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent
REPO = CSRC.parent.parent
INCLUDE = REPO / "include"
LIB = CSRC / "libtsfmx_b200.so"
OBJ_DIR = CSRC / "build"

SOURCES = [
    "common.cu",
    "preprocess.cu",
    "elementwise.cu",
    "attention.cu",
    "gemm.cu",
    "backward.cu",
    "attention_bwd.cu",
    "chronos.cu",
    "t5.cu",
    "decode.cu",
    "stack.cu",
]

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    "-Xcompiler",
    "-fPIC",
    f"-I{INCLUDE}",
    f"-I{CSRC}",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; tsfmx_b200 needs the CUDA toolkit to build its sm_100a kernels")


def _digest(paths: list[Path]) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(nvcc: str, src: Path, obj: Path, verbose: bool) -> str:
    cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    return r.stderr if verbose else ""


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / s for s in SOURCES]
    missing = [str(p) for p in srcs if not p.exists()]
    if missing:
        raise RuntimeError(f"missing CUDA sources: {missing}")
    deps = srcs + sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h"))
    stamp = OBJ_DIR / "stamp.txt"
    digest = _digest(deps)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    nvcc = _nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    objs = [OBJ_DIR / (s.stem + ".o") for s in srcs]
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as pool:
        futs = [pool.submit(_compile_one, nvcc, s, o, verbose) for s, o in zip(srcs, objs)]
        for f in futs:
            out = f.result()
            if out:
                print(out, file=sys.stderr)
    link = [nvcc, "-shared", "-o", str(LIB), *[str(o) for o in objs], "-cudart", "static"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
