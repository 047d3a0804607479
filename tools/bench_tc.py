"""Micro-benchmark of the tcgen05 kernels at the layer shapes of the 256x256 batch-32 step (CUDA events, each kernel timed
alone, 10 launches after 3 warm-ups): single-CTA conv (csrc/sg_conv_tc.cu), CTA-pair conv plain and with the fused
epilogue (csrc/sg_conv_pair_tc.cu), first- and second-generation weight gradient (csrc/sg_wgrad_tc.cu, sg_wgrad2_tc.cu)."""
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
hdr = ("N,HW,Cin,Cout", "conv1 ms", "TF/s", "pair ms", "TF/s", "pair+epi", "TF/s", "wgrad1 ms", "TF/s", "wgrad2 ms", "TF/s")
print("%-18s" % hdr[0] + "".join("%10s" % h for h in hdr[1:]))
nan = float("nan")
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
    ms_p = ms_pe = nan
    if lib.sg_conv3x3_pair_supported(N, R, R, ci, co):
        ms_p = t(lambda: lib.sg_conv3x3_pair(ptr(x), ptr(E), ptr(y), ptr(b), 1.0, 1, None, None, None, N, R, R, ci, co, st))
        noise = torch.randn(N, 1, R, R, device="cuda"); nw = torch.randn(co, device="cuda")
        part = torch.empty(N, lib.sg_conv3x3_pair_slots(N, R, R, ci, co), 2, co, device="cuda")
        ms_pe = t(lambda: lib.sg_conv3x3_pair(ptr(x), ptr(E), ptr(y), ptr(b), 1.0, 1, ptr(noise), ptr(nw), ptr(part), N, R, R,
                                              ci, co, st))
    ms_w = ms_w2 = nan
    if lib.sg_wgrad3x3_tc_supported(N, R, R, ci, co):
        S = lib.sg_wgrad3x3_tc_splits(N, R, R, ci, co)
        part1 = torch.empty(S, 9, co, ci, device="cuda")
        ms_w = t(lambda: lib.sg_wgrad3x3_tc(ptr(x), ptr(dy), ptr(part1), N, R, R, ci, co, st))
    if lib.sg_wgrad3x3_v2_supported(N, R, R, ci, co):
        S = lib.sg_wgrad3x3_v2_splits(N, R, R, ci, co)
        part2 = torch.empty(S, 9, co, ci, device="cuda")
        ms_w2 = t(lambda: lib.sg_wgrad3x3_v2(ptr(x), ptr(dy), ptr(part2), N, R, R, ci, co, st))
    row = [ms_c, ms_p, ms_pe, ms_w, ms_w2]
    print("%-18s" % ("%d,%d,%d,%d" % (N, R, ci, co)) + "".join("%10.4f%10.1f" % (m, flops / m / 1e9) for m in row))
