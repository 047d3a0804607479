// tcgen05 / TMA implicit-GEMM convolution (placeholder until the kernel lands in this file).
#include "sg_common.cuh"
#include "../../include/sgb200.h"

SG_EXPORT int sg_conv3x3_tc_supported(int N, int H, int W, int Cin, int Cout) { return 0; }

SG_EXPORT int sg_conv3x3_tc(const void* x, const void* Etc, void* y, const float* bias, float bias_mul, int act, int N,
                            int H, int W, int Cin, int Cout, void* stream) {
  sg_set_error("sg_conv3x3_tc: shape not supported");
  return 1002;
}
