"""CUDA-event timing of the TimesFM attention kernels at the benchmarked shapes (one B200):
forward at 4096 series x 16 patches (forecast), backward at 1024 series with and without the parameter gradients."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "multimodal-timesfm_b200"))
from tsfmx_b200 import ops  # noqa: E402


def timeit(fn, iters=50):
    for _ in range(5):
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
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    h, hd, n = 16, 80, 16
    inv_freq = (1.0 / (10000.0 ** (torch.arange(0, hd, 2, device=dev).float() / hd))).contiguous()
    qw, kw = (torch.rand(hd, device=dev, generator=g) + 0.5 for _ in range(2))
    qs = torch.rand(hd, device=dev, generator=g) * 0.2
    for b in (4096, 1024):
        rows = b * n
        # several buffers so that consecutive launches do not find their operands in L2
        qkvs = [torch.randn(rows, 3 * h * hd, device=dev, generator=g).bfloat16() for _ in range(4 if b == 4096 else 12)]
        pm = torch.zeros(b, n, dtype=torch.bool, device=dev)
        nm = torch.zeros(b, dtype=torch.int32, device=dev)
        out = torch.empty(rows, h * hd, dtype=torch.bfloat16, device=dev)
        i = [0]

        def fwd():
            i[0] += 1
            ops.timesfm_attention(qkvs[i[0] % len(qkvs)], b, n, h, hd, pm, nm, inv_freq, qw, kw, qs, 1e-6, ops.DT_BF16, out=out)

        t = timeit(fwd)
        print(f"forward  {b} series: {t:7.1f} us  ({rows * 8 * h * hd / t / 1e6:.2f} TB/s algorithmic)", flush=True)
        if b == 1024:
            dout = torch.randn(rows, h * hd, device=dev, generator=g).bfloat16()
            dqkv = torch.empty(rows, 3 * h * hd, dtype=torch.bfloat16, device=dev)
            dparams = torch.zeros(2 * hd, dtype=torch.float32, device=dev)
            for label, dp in (("backward", None), ("backward + parameter gradients", dparams)):
                def bwd():
                    i[0] += 1
                    ops.timesfm_attention_bwd(qkvs[i[0] % len(qkvs)], dout, b, n, h, hd, pm, nm, inv_freq, qw, kw, qs, 1e-6,
                                              ops.DT_BF16, dqkv=dqkv, dparams=dp)
                t = timeit(bwd)
                print(f"{label} {b} series: {t:7.1f} us  ({rows * 14 * h * hd / t / 1e6:.2f} TB/s algorithmic)", flush=True)


if __name__ == "__main__":
    main()
