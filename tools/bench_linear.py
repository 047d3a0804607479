"""Micro-benchmark of the dense-layer and weight-preparation entry points at the shapes of the 256x256 batch-32 step (CUDA
events around CUDA-graph replays of 20 launches, L2 warm as inside the step): mapping layers 512->512, the discriminator's dense0
8192->512 and dense1 512->1 (csrc/sg_small.cu), and sg_wprep_fwd / sg_wprep_bwd for the 3x3 weights
(csrc/sg_conv_generic.cu).  Prints microseconds and the effective weight-traffic bandwidth."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stylegan.pytorch_b200"))
import torch
from models._lib import lib, ptr, stream, code


def t(fn, iters=20):
    """microseconds per launch, measured over a CUDA-graph replay of `iters` launches (no host time between launches)."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * iters) * 1e3


print("%-34s%10s%10s%10s" % ("dense B,I,O", "fwd us", "dgrad us", "wgrad us"))
for B, I, O in [(32, 512, 512), (64, 512, 512), (32, 8192, 512), (32, 512, 1)]:
    x = torch.randn(B, I, device="cuda"); W = torch.randn(O, I, device="cuda"); b = torch.randn(O, device="cuda")
    y = torch.empty(B, O, device="cuda"); gy = torch.randn(B, O, device="cuda"); gx = torch.empty(B, I, device="cuda")
    gW = torch.empty(O, I, device="cuda"); gb = torch.empty(O, device="cuda")
    ws = torch.empty(max(1, lib.sg_linear_fwd_ws_floats(B, I, O)), device="cuda")
    f = t(lambda: lib.sg_linear_fwd(ptr(x), I, ptr(W), ptr(b), 0.1, 1.0, 1, ptr(y), B, I, O, ptr(ws), stream()))
    d = t(lambda: lib.sg_linear_dgrad(ptr(gy), ptr(W), 0.1, ptr(gx), B, I, O, stream()))
    w = t(lambda: lib.sg_linear_wgrad(ptr(gy), ptr(x), I, 0.1, 1.0, ptr(gW), ptr(gb), B, I, O, 0, stream()))
    print("%-34s%10.1f%10.1f%10.1f" % ("%d,%d,%d" % (B, I, O), f, d, w))

print("%-34s%10s%10s%10s%10s" % ("wprep co,ci,KK,npt,S", "fwd us", "GB/s", "bwd us", "GB/s"))
for co, ci, KK, npt, S in [(64, 64, 9, 9, 128), (128, 128, 9, 9, 32), (256, 256, 9, 9, 8), (512, 512, 9, 9, 3),
                           (512, 512, 9, 16, 3), (256, 512, 9, 16, 4)]:
    W = torch.randn(co, ci, KK, device="cuda"); coef = torch.randn(npt, KK, device="cuda")
    E = torch.empty(npt, co, ci, device="cuda", dtype=torch.bfloat16)
    part = torch.randn(S, npt, co, ci, device="cuda"); dW = torch.empty(co, ci, KK, device="cuda")
    for tr in (0, 1):
        f = t(lambda: lib.sg_wprep_fwd(ptr(W), ptr(E), code(torch.bfloat16), co, ci, KK, npt, ptr(coef), 0.1, tr, stream()))
        bw = t(lambda: lib.sg_wprep_bwd(ptr(part), S, ptr(dW), co, ci, KK, npt, ptr(coef), 0.1, tr, 0, stream()))
        fb = (W.numel() * 4 + E.numel() * 2) / f / 1e3
        bb = (part.numel() * 4 + dW.numel() * 4) / bw / 1e3
        print("%-34s%10.1f%10.1f%10.1f%10.1f" % ("%d,%d,%d,%d,%d t=%d" % (co, ci, KK, npt, S, tr), f, fb, bw, bb))
