"""Two eager single-stream forecasts of a 2-layer TimesFM stack at the benchmarked shape (4096 series, ctx 512: M = 65 536
token rows) - the program `ncu --set full` is pointed at to capture one decoder layer's seven launches
(qkv GEMM, attention, out GEMM, junction, ff0 GEMM, ff1 GEMM, junction) as the final tree issues them.  Not a bench line.

Launches matching `gemm_bf16_tcgen05|norm_residual_norm|timesfm_attention_mma` per forecast: tokenizer 2 + fusion 1 +
2 layers x 7 + head 2 = 19; layer 0 of the second forecast is launches 22..28 (--launch-skip 22 -c 7)."""
import os
import sys
import time

t0 = time.time()
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-timesfm_b200"))
sys.path.insert(0, ROOT)
from oracle import timesfm_oracle as O  # noqa: E402  synthetic inputs only
from tsfmx_b200 import _lib  # noqa: E402
from tsfmx_b200.decoder import MultimodalDecoder, MultimodalDecoderConfig  # noqa: E402
from tsfmx_b200.tsfm.timesfm import TimesFM2p5Adapter, init_random_  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
adapter = TimesFM2p5Adapter(num_layers=2, precision="bf16", with_quantile_head=False)
adapter.stack_call = False  # per-kernel entry points: same kernels, same order as the whole-stack call
init_random_(adapter, seed=0)
dec = MultimodalDecoder(adapter, MultimodalDecoderConfig(384, 1, [])).to(dev).eval()
dec.set_precision("bf16")
dec.lanes = 1
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx, masks, text, _ = O.synthetic_batch(B, 512, 128, seed=1234)
ctx, masks, text = ctx.to(dev), masks.to(dev), text.to(dev)
with torch.no_grad():
    for i in range(2):
        n0 = _lib.launch_count()
        y = dec(128, ctx, masks, text)
        torch.cuda.synchronize()
        print(f"forecast {i}: {_lib.launch_count() - n0} launches, |y| max {y.abs().max().item():.4f}, t = {time.time() - t0:.1f} s",
              flush=True)
