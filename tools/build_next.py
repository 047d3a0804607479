"""Builds the next-round candidate kernels (stylegan.pytorch_b200/csrc_next/*.cu) into lib/libsgb200_next.so.
They are NOT part of the product library; tests/test_next_kernels.py (-m gpu_next) and tools/bench_next.py load this
library with ctypes and compare against libsgb200.so.   python tools/build_next.py"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "stylegan.pytorch_b200", "csrc_next")
OUT = os.path.join(ROOT, "stylegan.pytorch_b200", "lib", "libsgb200_next.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(SRC, "*.cu")))
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "-shared", "-cudart", "static", "-o", OUT] + srcs
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build())
