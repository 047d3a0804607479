"""Runs the north-star kernel (fused 3x3 conv 64->64 @256x256, N=32, bf16) a few times; target of `ncu --set full`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
print(bench.northstar_kernel_roofline(32, 256, 64, iters=n))
