"""Probe: split-K GEMM numerics on the Chronos-2 FFN shapes (K = 3072) with poisoned output buffers, against fp64
matmul, for split modes never / auto / forced 3.  Printed worst relative errors were 1e-6..7e-6 in every mode (the
split sums are slightly MORE accurate); the run-to-run variation of a gradient test came from the reduction order,
which is why automatic splitting starts at K = 4096 (see csrc/gemm.cu).  Not a bench line."""
import sys, torch
sys.path.insert(0, "/root/repo/multimodal-timesfm_b200"); sys.path.insert(0, "/root/repo")
from tsfmx_b200 import _lib, ops
from tsfmx_b200._lib import DT_F32, DT_BF16_SPLIT, PREC_BF16X3, PREC_BF16, DT_BF16
lib = _lib.load()
dev = "cuda"
torch.manual_seed(0)
junk = [torch.full((64, 1024, 1024), float("nan"), device=dev) for _ in range(2)]  # poison the allocator
del junk
for (m, n, k) in [(450, 768, 3072), (450, 768, 3072), (225, 768, 3072), (450, 3072, 768), (1350, 768, 3072)]:
    for prec in (PREC_BF16X3, PREC_BF16):
        x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev) * 0.05
        if prec == PREC_BF16X3:
            a, b = ops.cast_rows(x, DT_BF16_SPLIT), ops.cast_rows(w, DT_BF16_SPLIT)
            ref = (x.double() @ w.double().t()).float()
        else:
            a, b = x.bfloat16(), w.bfloat16()
            ref = (a.double() @ b.double().t()).float()
        res = {}
        for mode in (1, 0, 3):
            _lib.check(lib.tsfmx_gemm_set_split_k(mode))
            worst = 0.0
            for it in range(20):
                out = torch.full((m, n), float("nan"), device=dev)
                ops.gemm([(a, b, k)], m, n, out, DT_F32, precision=prec)
                worst = max(worst, ((out - ref).abs().max() / ref.abs().max()).item())
            res[mode] = worst
        print((m, n, k), "x3" if prec == PREC_BF16X3 else "bf16", {k_: f"{v:.2e}" for k_, v in res.items()}, flush=True)
_lib.check(lib.tsfmx_gemm_set_split_k(0))
