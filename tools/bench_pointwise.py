"""CUDA-event timings of the bandwidth-bound kernels at the 256x256 / 64-channel / batch-32 shape of the step (each
kernel alone, 3 warm-ups + `iters` timed launches), as algorithmic GB/s.  Also the target of
`ncu --set full -k regex:<kernel>` (pass `--iters 1 --only <name>`).   python tools/bench_pointwise.py"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stylegan.pytorch_b200"))
import torch
from models._lib import lib, ptr, stream, code

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--only", default="")
ap.add_argument("--res", type=int, default=256)
ap.add_argument("--ch", type=int, default=64)
ap.add_argument("--batch", type=int, default=32)
a = ap.parse_args()
dev = torch.device("cuda", 0)
N, H, C = a.batch, a.res, a.ch
HW, P = H * H, a.batch * H * H
bf = torch.bfloat16
BF = code(bf)
x = torch.randn(N, H, H, C, device=dev).to(bf)
x2 = torch.randn(N, H, H, C, device=dev).to(bf)
y = torch.empty_like(x)
img = torch.randn(N, 3, H, H, device=dev)
img_out = torch.empty_like(img)
W3 = torch.randn(3, C, device=dev)
Wc3 = torch.randn(C, 3, device=dev)
b3, bC = torch.randn(3, device=dev), torch.randn(C, device=dev)
coef = torch.randn(N, C, 4, device=dev)
ab = torch.randn(N, C, 3, device=dev)
wt = torch.randn(N, 1, H, H, device=dev)
nblk = lib.sg_channel_sum_blocks(P)
ws = torch.empty(nblk * 2 * C, device=dev)
rows = torch.empty(2, C, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > L2 (126 MB)
T = x.numel() * 2                                                       # bytes of one bf16 feature map
S = stream()

CASES = {
    "lrelu_bwd": (3 * T, lambda: lib.sg_lrelu_bwd(ptr(x), ptr(x2), ptr(y), x.numel(), BF, S)),
    "axpby": (3 * T, lambda: lib.sg_axpby(ptr(x), ptr(x2), ptr(y), 0.5, 0.5, x.numel(), BF, S)),
    "blur3": (2 * T, lambda: lib.sg_blur3(ptr(x), ptr(y), None, 1.0, BF, N, H, H, C, S)),
    "epi_bwd_apply": (3 * T, lambda: lib.sg_epi_bwd_apply(ptr(x), ptr(x2), ptr(coef), 1, ptr(y), BF, N, HW, C, S)),
    "affine_nc": (2 * T, lambda: lib.sg_affine_nc(ptr(x), ptr(ab), ptr(y), BF, N, HW, C, S)),
    "rgb_out": (T + img.numel() * 4, lambda: lib.sg_rgb_out(ptr(x), BF, ptr(W3), ptr(b3), 0.1, 1.0, 1.0, 0.0,
                                                            ptr(img_out), N, H, H, C, 3, 1, S)),
    "rgb_out_dgrad": (T + img.numel() * 4, lambda: lib.sg_rgb_out_dgrad(ptr(img), ptr(W3), 0.1, ptr(y), BF, N, H, H,
                                                                        C, 3, 1, S)),
    "rgb_in": (T + img.numel() * 4, lambda: lib.sg_rgb_in(ptr(img), ptr(Wc3), ptr(bC), 0.1, 1.0, 1, ptr(y), BF, N, HW,
                                                          3, C, S)),
    "rgb_in_dgrad": (T + img.numel() * 4, lambda: lib.sg_rgb_in_dgrad(ptr(x), BF, ptr(Wc3), 0.1, ptr(img_out), N, HW,
                                                                      3, C, S)),
    "channel_sum_w": (T, lambda: lib.sg_channel_sum(ptr(x), BF, ptr(wt), 1, 1, P, HW, C, ptr(ws), ptr(rows), 1.0, 0,
                                                    S)),
    "channel_sum_ones": (T, lambda: lib.sg_channel_sum(ptr(x), BF, None, 0, 1, P, HW, C, ptr(ws), ptr(rows), 1.0, 0,
                                                       S)),
}

print("%-18s %10s %10s   (N=%d, %dx%d, C=%d, bf16; L2 flushed between launches)" % ("kernel", "us", "GB/s", N, H, H, C))
for name, (nbytes, fn) in CASES.items():
    if a.only and a.only != name:
        continue
    for _ in range(0 if a.iters == 1 else 3):
        fn()                                   # the binding raises on a non-zero status
    tot = 0.0
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = tot / a.iters * 1e3
    print("%-18s %10.1f %10.0f" % (name, us, nbytes / us * 1e-3))
