"""HBM-bound stages at sizes >> L2 (B >= 262144 series at ctx 512): achieved GB/s from the
ALGORITHMIC bytes and CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.

    python scripts/bench_hbm_kernels.py [--iters 10] [--only NAME]      -> one JSON line per kernel
"""

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "multimodal-timesfm_b200"))

import torch  # noqa: E402

from tsfmx_b200 import ops  # noqa: E402
from tsfmx_b200._lib import DT_BF16, DT_F32  # noqa: E402


def peak_gbs() -> float:
    p = ROOT / "MEASURED_PEAKS.json"
    return float(json.loads(p.read_text())["hbm_gbs"]) if p.exists() else 6650.0


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def stage_cases(dev, contexts):
    """(name, algorithmic bytes per series, series, launch) for the three HBM-bound stages at every (context, batch)."""
    cases = []
    for ctx_len, batch in contexts:
        g = torch.Generator(device=dev).manual_seed(ctx_len)
        x = torch.randn(batch, ctx_len, generator=g, device=dev)
        mask = torch.zeros(batch, ctx_len, dtype=torch.bool, device=dev)
        n32, n16 = ctx_len // 32, ctx_len // 16
        cases += [
            (f"timesfm_patchify_norm f32-out ctx{ctx_len}", 5 * ctx_len + 2 * ctx_len * 4 + 9 * n32, batch,
             lambda x=x, mask=mask: ops.timesfm_patchify_norm(x, mask, 32, DT_F32)),
            (f"timesfm_patchify_norm bf16-out ctx{ctx_len}", 5 * ctx_len + 2 * ctx_len * 2 + 9 * n32, batch,
             lambda x=x, mask=mask: ops.timesfm_patchify_norm(x, mask, 32, DT_BF16)),
            (f"chronos2_patchify_norm f32-out ctx{ctx_len}", 5 * ctx_len + 3 * ctx_len * 4 + n16 + 8, batch,
             lambda x=x, mask=mask: ops.chronos2_patchify_norm(x, mask, 16, True, 8192.0, DT_F32)),
            (f"chronos2_patchify_norm bf16-out ctx{ctx_len}", 5 * ctx_len + 3 * ctx_len * 2 + n16 + 8, batch,
             lambda x=x, mask=mask: ops.chronos2_patchify_norm(x, mask, 16, True, 8192.0, DT_BF16)),
        ]
        centers = torch.linspace(-15.0, 15.0, 4093)
        bounds = torch.cat([torch.tensor([-1e20]), (centers[1:] + centers[:-1]) / 2, torch.tensor([1e20])]).to(dev)
        cases.append((f"chronos_t5_tokenize int64-ids ctx{ctx_len}", 4 * ctx_len + 9 * (ctx_len + 1) + 4, batch,
                      lambda x=x, bounds=bounds: ops.chronos_t5_tokenize(x, bounds)))
    return cases


def measure(cases, iters, peak, only=""):
    """One record per case: achieved GB/s = algorithmic bytes x series / CUDA-event time per launch."""
    out = []
    for name, bytes_per_series, batch, fn in cases:
        if only and only not in name:
            continue
        ms = timeit(fn, iters)
        gbs = bytes_per_series * batch / (ms * 1e-3) / 1e9
        out.append({
            "kernel": name, "bound": "hbm", "series": batch, "algorithmic_bytes_per_series": bytes_per_series,
            "ms": round(ms, 4), "series_per_s": batch / (ms * 1e-3), "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s",
            "frac": round(gbs / peak, 3),
            "note": "time includes torch.empty of the outputs; working set >> 126 MB L2",
        })
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--contexts", default="", help="comma-separated context lengths (default 512,2048)")
    ap.add_argument("--tune", default="", help="comma-separated knob=value pairs passed to tsfmx_tune (kernel A/B)")
    args = ap.parse_args()
    from tsfmx_b200 import _lib

    for pair in filter(None, args.tune.split(",")):
        knob, value = pair.split("=")
        _lib.check(_lib.load().tsfmx_tune(int(knob), int(value)))
    dev = torch.device("cuda")
    contexts = [(512, 262144), (2048, 65536)] if not args.contexts else [
        (int(c), max(8192, 262144 * 512 // int(c))) for c in args.contexts.split(",")]
    for rec in measure(stage_cases(dev, contexts), args.iters, peak_gbs(), args.only):
        rec["tune"] = args.tune
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
