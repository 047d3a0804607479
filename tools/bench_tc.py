"""Micro-benchmark of the tcgen05 kernels at the layer shapes of the 256x256 batch-32 step (CUDA events, timed alone)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stylegan.pytorch_b200"))
import torch
from models import _tc
from models._lib import lib, ptr, stream

def t(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

shapes = [(32, 256, 64, 64), (32, 256, 128, 64), (32, 256, 64, 128), (32, 128, 128, 128), (32, 128, 256, 128),
          (32, 64, 256, 256), (32, 64, 512, 256), (32, 32, 512, 512), (32, 16, 512, 512)]
print("%-28s %10s %10s %10s %10s" % ("N,HW,Cin,Cout", "conv ms", "conv TF/s", "wgrad ms", "wgrad TF/s"))
for N, R, ci, co in shapes:
    x = torch.randn(N, R, R, ci, device="cuda").bfloat16()
    dy = torch.randn(N, R, R, co, device="cuda").bfloat16()
    W = torch.nn.Parameter(torch.randn(co, ci, 3, 3, device="cuda"))
    b = torch.randn(co, device="cuda")
    y = torch.empty(N, R, R, co, device="cuda", dtype=torch.bfloat16)
    E = _tc.tc_weight(W, "fwd", 0.05)
    st = stream()
    flops = 2.0 * N * R * R * ci * co * 9
    ms_c = t(lambda: lib.sg_conv3x3_tc(ptr(x), ptr(E), ptr(y), ptr(b), 1.0, 1, N, R, R, ci, co, st))
    ms_w = float("nan")
    if lib.sg_wgrad3x3_tc_supported(N, R, R, ci, co):
        S = lib.sg_wgrad3x3_tc_splits(N, R, R, ci, co)
        part = torch.empty(S, 9, co, ci, device="cuda")
        ms_w = t(lambda: lib.sg_wgrad3x3_tc(ptr(x), ptr(dy), ptr(part), N, R, R, ci, co, st))
    print("%-28s %10.4f %10.1f %10.4f %10.1f" % ("%d,%d,%d,%d" % (N, R, ci, co), ms_c, flops / ms_c / 1e9, ms_w, flops / ms_w / 1e9))
