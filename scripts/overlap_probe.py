"""Does a tensor-bound GEMM co-run with an HBM-bound kernel launched on another stream?  Prints alone / together times."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "multimodal-timesfm_b200"))
import torch

from tsfmx_b200 import ops
from tsfmx_b200._lib import DT_BF16

dev = torch.device("cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = 1280
torch.manual_seed(0)
xn = torch.randn(M, D, device=dev).bfloat16()
w_qkv = (0.02 * torch.randn(3 * D, D, device=dev)).bfloat16()
w_o = (0.02 * torch.randn(D, D, device=dev)).bfloat16()
qkv = torch.empty(M, 3 * D, dtype=torch.bfloat16, device=dev)
o = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
a = torch.randn(M, D, device=dev).bfloat16()
x = torch.randn(M, D, device=dev)
y = torch.empty_like(x)
yn = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
wp = torch.ones(D, device=dev)
qkv2 = torch.randn(M, 3 * D, device=dev).bfloat16()
att = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
B, N = M // 16, 16
pm = torch.zeros(B, N, dtype=torch.bool, device=dev)
nm = torch.zeros(B, dtype=torch.int32, device=dev)
inv_freq = (1.0 / (10000.0 ** (torch.arange(0, 80, 2).float() / 80))).to(dev)
hw = torch.ones(80, device=dev)


def k_qkv():
    ops.gemm([(xn, w_qkv, D)], M, 3 * D, qkv, DT_BF16)


def k_out():
    ops.gemm([(xn, w_o, D)], M, D, o, DT_BF16)


def k_nrn():
    ops.norm_residual_norm(a, x, wp, wp, 1e-6, y, DT_BF16, yn)


def k_att():
    ops.timesfm_attention(qkv2, B, N, 16, 80, pm, nm, inv_freq, hw, hw, hw, 1e-6, DT_BF16, out=att)


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def alone(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def together(f1, n1, f2, n2, reps=10):
    """n1 launches of f1 on s1 and n2 of f2 on s2 per rep; returns us per rep (wall, device) and per-stream spans."""
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record()
    s1.wait_stream(main)
    s2.wait_stream(main)
    with torch.cuda.stream(s1):
        a0.record()
    with torch.cuda.stream(s2):
        b0.record()
    for _ in range(reps):
        with torch.cuda.stream(s1):
            for _ in range(n1):
                f1()
        with torch.cuda.stream(s2):
            for _ in range(n2):
                f2()
    with torch.cuda.stream(s1):
        a1.record()
    with torch.cuda.stream(s2):
        b1.record()
    main.wait_stream(s1)
    main.wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, a0.elapsed_time(a1) / reps * 1e3, b0.elapsed_time(b1) / reps * 1e3


t = {n: alone(f) for n, f in [("qkv", k_qkv), ("out", k_out), ("nrn", k_nrn), ("att", k_att)]}
print("alone us:", {k: round(v, 1) for k, v in t.items()}, "M =", M)
for (n1, f1, c1), (n2, f2, c2) in [
    (("qkv", k_qkv, 1), ("nrn", k_nrn, 2)),
    (("qkv", k_qkv, 1), ("att", k_att, 2)),
    (("out", k_out, 2), ("nrn", k_nrn, 2)),
    (("qkv", k_qkv, 1), ("out", k_out, 3)),
    (("nrn", k_nrn, 1), ("att", k_att, 1)),
]:
    tot, sa, sb = together(f1, c1, f2, c2)
    print(f"{c1}x{n1} || {c2}x{n2}: together {tot:.1f} us (stream spans {sa:.1f} / {sb:.1f}); serial sum {c1 * t[n1] + c2 * t[n2]:.1f} us")

# ---- power / clocks under sustained load of each kind (is the device power-capped, so that overlap cannot pay?)
import subprocess, threading, time, statistics


def sample_power(stop, out):
    while not stop.is_set():
        r = subprocess.run(["nvidia-smi", "--id=0", "--query-gpu=power.draw,clocks.sm,power.limit,temperature.gpu",
                            "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
        try:
            out.append([float(v) for v in r.split(",")])
        except ValueError:
            pass
        stop.wait(0.1)


def sustained(name, body, seconds=2.5):
    stop, out = threading.Event(), []
    th = threading.Thread(target=sample_power, args=(stop, out))
    torch.cuda.synchronize()
    th.start()
    t0 = time.time()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(20):
            body()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    tail = out[len(out) // 2:]
    print(f"{name}: {e0.elapsed_time(e1) / n * 1e3:.1f} us/iter, power {statistics.median(v[0] for v in tail):.0f} W "
          f"(limit {tail[-1][2]:.0f}), sm {statistics.median(v[1] for v in tail):.0f} MHz, temp {tail[-1][3]:.0f} C")


def both():
    main = torch.cuda.current_stream()
    s1.wait_stream(main); s2.wait_stream(main)
    with torch.cuda.stream(s1):
        k_qkv()
    with torch.cuda.stream(s2):
        k_nrn(); k_nrn()
    main.wait_stream(s1); main.wait_stream(s2)


sustained("qkv only", k_qkv)
sustained("nrn only", k_nrn)
sustained("att only", k_att)
sustained("qkv || 2 nrn", both)
