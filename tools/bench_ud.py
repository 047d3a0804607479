"""Micro-benchmark of the up/down-sampling tensor-core convolution (sg_convud_tc) at the layer shapes of the 256x256
batch-32 step, for each kernel generation (sg_pair_updown: 2 halo-box pair kernel csrc/sg_convud2_tc.cu, 1 first pair
kernel, 0 single CTA; csrc/sg_convud_tc.cu).  CUDA events, each kernel alone, 10 launches after 3 warm-ups.
FLOPs = 16 tap-GEMMs per low-resolution pixel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stylegan.pytorch_b200"))
import torch
from models._lib import lib, ptr, stream


def t(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


UP, DOWN = 0, 1
shapes = [(UP, 32, 128, 128, 64), (UP, 32, 64, 256, 128), (UP, 32, 32, 512, 256), (UP, 32, 16, 512, 512), (UP, 32, 8, 512, 512),
          (DOWN, 32, 128, 64, 128), (DOWN, 32, 64, 128, 256), (DOWN, 32, 32, 256, 512), (DOWN, 32, 16, 512, 512),
          (UP, 32, 128, 128, 64), (UP, 32, 128, 64, 64), (DOWN, 32, 128, 64, 64)]
print("%-26s" % "mode,N,lo,Cin,Cout" + "".join("%12s%9s" % ("gen%d ms" % g, "TF/s") for g in (2, 1, 0)))
for mode, N, lo, ci, co in shapes:
    hin = lo if mode == UP else 2 * lo
    hout = 2 * lo if mode == UP else lo
    x = torch.randn(N, hin, hin, ci, device="cuda").bfloat16()
    E = (torch.randn(16, co, ci, device="cuda") * 0.05).bfloat16()
    b = torch.randn(co, device="cuda")
    y = torch.empty(N, hout, hout, co, device="cuda", dtype=torch.bfloat16)
    flops = 2.0 * N * lo * lo * ci * co * 16
    row = []
    for gen in (2, 1, 0):
        if not lib.sg_convud_tc_supported(mode, N, lo, lo, ci, co):
            row.append(float("nan")); continue
        old = lib.sg_pair_updown(gen)
        try:
            row.append(t(lambda: lib.sg_convud_tc(mode, ptr(x), ptr(E), ptr(y), ptr(b), 1.0, 1, N, lo, lo, ci, co, stream())))
        finally:
            lib.sg_pair_updown(old)
    print("%-26s" % ("%s,%d,%d,%d,%d" % ("UP" if mode == UP else "DOWN", N, lo, ci, co)) +
          "".join("%12.4f%9.1f" % (m, flops / m / 1e9) for m in row))
