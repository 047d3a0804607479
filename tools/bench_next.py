"""Product kernel vs next-round candidate (lib/libsgb200_next.so, tools/build_next.py) at the step's 256x256x64 batch-32
shape: CUDA-event time of each alone, L2 flushed between launches.   python tools/bench_next.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stylegan.pytorch_b200"))
import torch
from models._lib import lib, ptr, stream

nxt = ctypes.CDLL(os.path.join(ROOT, "stylegan.pytorch_b200", "lib", "libsgb200_next.so"))
vp, ci, cf, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong
nxt.sgn_blur3.argtypes = [vp, vp, vp, cf, ci, ci, ci, ci, ci, vp]
nxt.sgn_rgb_in3.argtypes = [vp, vp, vp, cf, cf, ci, vp, ci, ci, ci, vp]
nxt.sgn_rgb_out3.argtypes = [vp, vp, vp, cf, cf, cf, cf, vp, ci, ci, ci, ci, ci, vp]
nxt.sgn_channel_sum_blocks.argtypes = [cll]
nxt.sgn_channel_sum_w1.argtypes = [vp, vp, cll, ci, vp, vp, cf, vp]

dev = torch.device("cuda", 0)
N, H, C = 32, 256, 64
HW, P = H * H, N * H * H
x = torch.randn(N, H, H, C, device=dev).to(torch.bfloat16)
y = torch.empty_like(x)
img = torch.randn(N, 3, H, H, device=dev)
img_out = torch.empty_like(img)
W3, Wc3, b3, bC = torch.randn(3, C, device=dev), torch.randn(C, 3, device=dev), torch.randn(3, device=dev), torch.randn(C, device=dev)
wt = torch.randn(N, 1, H, H, device=dev)
ws = torch.empty(lib.sg_channel_sum_blocks(P) * 2 * C, device=dev)
rows = torch.empty(2, C, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
S = stream()
T = x.numel() * 2
I = img.numel() * 4


def chk(rc):
    assert rc in (0, None), rc


CASES = [
    ("blur3", 2 * T, lambda: lib.sg_blur3(ptr(x), ptr(y), None, 1.0, 1, N, H, H, C, S),
     lambda: chk(nxt.sgn_blur3(x.data_ptr(), y.data_ptr(), None, 1.0, 1, N, H, H, C, S))),
    ("rgb_in", T + I, lambda: lib.sg_rgb_in(ptr(img), ptr(Wc3), ptr(bC), 0.1, 1.0, 1, ptr(y), 1, N, HW, 3, C, S),
     lambda: chk(nxt.sgn_rgb_in3(img.data_ptr(), Wc3.data_ptr(), bC.data_ptr(), 0.1, 1.0, 1, y.data_ptr(), N, HW, C, S))),
    ("rgb_out", T + I, lambda: lib.sg_rgb_out(ptr(x), 1, ptr(W3), ptr(b3), 0.1, 1.0, 1.0, 0.0, ptr(img_out), N, H, H, C, 3, 1, S),
     lambda: chk(nxt.sgn_rgb_out3(x.data_ptr(), W3.data_ptr(), b3.data_ptr(), 0.1, 1.0, 1.0, 0.0, img_out.data_ptr(), N, H, H,
                                  C, 1, S))),
    ("channel_sum_w", T, lambda: lib.sg_channel_sum(ptr(x), 1, ptr(wt), 1, 1, P, HW, C, ptr(ws), ptr(rows), 1.0, 0, S),
     lambda: chk(nxt.sgn_channel_sum_w1(x.data_ptr(), wt.data_ptr(), P, C, ws.data_ptr(), rows.data_ptr(), 1.0, S))),
]


def time_it(fn, iters=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


print("%-16s %12s %12s   (us, GB/s; N=%d %dx%d C=%d bf16)" % ("kernel", "product", "candidate", N, H, H, C))
for name, nbytes, prod, cand in CASES:
    a, b = time_it(prod), time_it(cand)
    print("%-16s %6.1f %5.0f %6.1f %5.0f" % (name, a, nbytes / a * 1e-3, b, nbytes / b * 1e-3))
