"""Profiling driver (GPU box): one warm-up G+D step, then ONE step inside cudaProfilerStart/Stop.
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/profile_step.py [--res 256 --batch 32 --precision bf16]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=256)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--warmup", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
sg = bench.make_gan(a.res, a.precision, dev, 1234)
depth = bench.O_depth(a.res)
img = torch.randn(a.batch, 3, a.res, a.res, device=dev)
z = torch.randn(a.batch, 512, device=dev)
for _ in range(a.warmup):
    sg.optimize_discriminator(z, img, depth, 0.5); sg.optimize_generator(z, img, depth, 0.5)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
sg.optimize_discriminator(z, img, depth, 0.5)
sg.optimize_generator(z, img, depth, 0.5)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
