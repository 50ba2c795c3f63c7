"""Probe: where does the end-to-end (host buffers -> MultimodalEvaluator.evaluate) step lose time against the
device-resident forward?  Prints ms/step for: forward only; evaluate over device-resident batches (no H2D);
evaluate over pinned host batches (the bench's e2e); the bare H2D copy.  Not a bench line."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-timesfm_b200"))
sys.path.insert(0, ROOT)
from oracle import timesfm_oracle as O  # noqa: E402  synthetic inputs only
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.evaluator import MultimodalEvaluator  # noqa: E402
from tsfmx_b200.tsfm.timesfm import TimesFM2p5Adapter, init_random_  # noqa: E402

dev = torch.device("cuda", 0)
adapter = TimesFM2p5Adapter(num_layers=50, precision="bf16", with_quantile_head=False)
init_random_(adapter, seed=0)
dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, [])).to(dev).eval()
dec.set_precision("bf16")
dec.lanes = int(os.environ.get("LANES", "2"))
B, STEPS = 4096, 10
host = []
for i in range(2):
    ctx, masks, text, hor = O.synthetic_batch(B, 512, 128, seed=1234 + i)
    host.append({"context": ctx.pin_memory(), "horizon": hor.pin_memory(), "text_embeddings": text.pin_memory()})
resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
masks = torch.zeros(B, 512, dtype=torch.bool, device=dev)
ev = MultimodalEvaluator(dec, dev)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / STEPS)
    return best


def fwd():
    for i in range(STEPS):
        b = resident[i % 2]
        dec(128, b["context"], masks, b["text_embeddings"])


def h2d():
    for i in range(STEPS):
        for k, v in host[i % 2].items():
            resident[i % 2][k].copy_(v, non_blocking=True)


for _ in range(2):
    fwd()
print(f"forward only                     : {timed(fwd):7.2f} ms/step")
print(f"evaluate, device-resident batches: {timed(lambda: ev.evaluate(resident[i % 2] for i in range(STEPS))):7.2f} ms/step")
print(f"evaluate, pinned host batches    : {timed(lambda: ev.evaluate(host[i % 2] for i in range(STEPS))):7.2f} ms/step")
print(f"bare H2D of one batch            : {timed(h2d):7.2f} ms/step "
      f"({sum(v.numel() * v.element_size() for v in host[0].values()) / 1e6:.0f} MB)")
