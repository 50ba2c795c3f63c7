"""Token-major weight-gradient GEMM (tsfmx_gemm_wgrad) against torch fp32, and against the transposing path it
replaces: python scripts/wgrad_probe.py  (one B200)."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "multimodal-timesfm_b200"))
from tsfmx_b200 import ops  # noqa: E402

CASES = [(4096 + 37, 1280, 1280), (16384, 3840, 1280), (1000, 336, 768), (513, 64, 1280), (200, 1280, 64), (65536, 1280, 1280)]


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def old_path(dy, x, rows, n_out, k_in):
    dy_t, kpad = ops.transpose_mask(dy, rows, n_out, ops.DT_BF16)
    x_t, _ = ops.transpose_mask(x, rows, k_in, ops.DT_BF16)
    gw = torch.empty(n_out, k_in, dtype=torch.float32, device=dy.device)
    ops.gemm([(dy_t, x_t, kpad)], n_out, k_in, gw, ops.DT_F32, precision=ops.PREC_BF16)
    return gw


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    argparse.ArgumentParser(description=__doc__).parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    for rows, n_out, k_in in CASES:
        dy = torch.randn(rows, n_out, device=dev, generator=g).bfloat16()
        x = torch.randn(rows, k_in, device=dev, generator=g).bfloat16()
        ref = dy.float().t() @ x.float()
        got = ops.wgrad(dy, x, rows, n_out, k_in, ops.PREC_BF16)
        torch.cuda.synchronize()
        print(f"rows {rows} n_out {n_out} k_in {k_in}: rel_max {rel(got, ref):.3e}"
              f"  (transposing path {rel(old_path(dy, x, rows, n_out, k_in), ref):.3e})", flush=True)
    for rows, n_out, k_in in [(16384, 1280, 1280), (16384, 3840, 1280), (65536, 1280, 1280)]:
        dy = torch.randn(rows, n_out, device=dev, generator=g).bfloat16()
        x = torch.randn(rows, k_in, device=dev, generator=g).bfloat16()
        t_new = timeit(lambda: ops.wgrad(dy, x, rows, n_out, k_in, ops.PREC_BF16))
        t_old = timeit(lambda: old_path(dy, x, rows, n_out, k_in))
        fl = 2.0 * rows * n_out * k_in
        print(f"rows {rows} n_out {n_out} k_in {k_in}: token-major {t_new:.1f} us ({fl / t_new / 1e6:.0f} TFLOP/s), "
              f"transposing {t_old:.1f} us ({fl / t_old / 1e6:.0f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    main()
