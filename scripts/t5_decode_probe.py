"""Probe: where does a Chronos-T5 greedy decode step spend its time?  Runs the eager decode loop for a few tokens over a
2048-series encoder output (random states: only the shapes matter) - under `ncu --metrics gpu__time_duration.sum` this
gives the per-kernel budget of a step.  Not a bench line."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT / "multimodal-timesfm_b200", ROOT):
    sys.path.insert(0, str(p))

import torch  # noqa: E402

from tsfmx_b200.tsfm import chronos_t5 as CT5  # noqa: E402


def main():
    series = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    adapter = CT5.ChronosT5Adapter(CT5.ChronosT5Module(), precision="bf16")
    CT5.init_random_(adapter._model, seed=0)
    adapter = adapter.cuda().eval()
    enc = torch.randn(series, 513, 768, device="cuda") * 0.1
    mask = torch.ones(series, 513, dtype=torch.bool, device="cuda")
    with torch.no_grad():
        adapter._decode_eager(enc, mask, 2, None, False)  # warm: packs, function attributes
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        adapter._decode_eager(enc, mask, steps, None, False)
        e1.record()
        torch.cuda.synchronize()
    print(f"{series} series, {steps} steps (+ cross K/V projection): {e0.elapsed_time(e1):.1f} ms")


if __name__ == "__main__":
    main()
