"""Probe: does running the 4096-series batch as sequential chunks (each chunk through all layers before the next)
keep the residual stream L2-resident and pay?  Prints series/s per chunk size.  Not a bench line.

Result on one B200 (round 1): no.  With the host cost removed by CUDA graphs: 4096 series per chunk 48.3 k series/s,
1024: 45.9 k, 512: 41.6 k, 256: 35.2 k - the GEMMs lose more at small M (wave quantisation, prologue/epilogue not
amortised) than the norm junctions gain from L2."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-timesfm_b200"))
sys.path.insert(0, ROOT)
from oracle import timesfm_oracle as O  # noqa: E402  synthetic inputs only
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.tsfm.timesfm import TimesFM2p5Adapter, init_random_  # noqa: E402

dev = torch.device("cuda", 0)
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 50
adapter = TimesFM2p5Adapter(num_layers=layers, precision="bf16", with_quantile_head=False)
init_random_(adapter, seed=0)
dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, [])).to(dev).eval()
dec.set_precision("bf16")
B = 4096
ctx, masks, text, _ = O.synthetic_batch(B, 512, 128, seed=1234)
ctx, masks, text = ctx.to(dev), masks.to(dev), text.to(dev)


def run(chunk, lanes):
    dec.lanes = lanes
    outs = []
    for s in range(0, B, chunk):
        outs.append(dec(128, ctx[s:s + chunk], masks[s:s + chunk], text[s:s + chunk]))
    return outs


def graph_run(chunk):
    """One CUDA graph of a chunk's whole forward, replayed per chunk (removes the host launch cost from the picture)."""
    dec.lanes = 1
    sc, sm, st = ctx[:chunk].clone(), masks[:chunk].clone(), text[:chunk].clone()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            dec(128, sc, sm, st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        out = dec(128, sc, sm, st)

    def go():
        for s0 in range(0, B - chunk + 1, chunk):
            sc.copy_(ctx[s0:s0 + chunk]); sm.copy_(masks[s0:s0 + chunk]); st.copy_(text[s0:s0 + chunk])
            g.replay()
    return go, out


import time
for chunk in (4096, 1024, 896, 512, 448, 256, 224):
    go, _ = graph_run(chunk)
    n = (B // chunk) * chunk
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        go()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    dec.lanes = 1
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dec(128, ctx[:chunk], masks[:chunk], text[:chunk])
    host_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    print(f"graph chunk={chunk:5d}: {ms:8.2f} ms/step {n / ms * 1e3:9.0f} series/s   (eager enqueue of one chunk: {host_ms:.1f} ms host)", flush=True)
    del go

# the same chunking with eager launches (host-bound below ~512 series per chunk: ~13 ms of Python per forward)
for chunk, lanes in ((4096, 2), (4096, 1), (2048, 1), (1024, 1), (896, 1), (448, 1), (512, 1), (224, 1), (256, 1), (448, 2), (896, 2)):
    for _ in range(3):
        run(chunk, lanes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = 5
    for _ in range(steps):
        run(chunk, lanes)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"chunk={chunk:5d} series ({chunk * 16:6d} tokens) lanes={lanes}: {ms:8.2f} ms/step  {B / ms * 1e3:9.0f} series/s", flush=True)
