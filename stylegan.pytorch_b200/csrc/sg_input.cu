// Input pipeline on the GPU: the per-image work of the reference's torchvision transform chain
// RandomHorizontalFlip -> ToTensor -> Normalize(0.5, 0.5)  (data/transforms.py:11-33) on a uint8 batch that was copied
// to the device as decoded (HWC bytes, 4x smaller than the float tensors the reference ships through its DataLoader).
//   out[n,k,h,w] = (u8[n,h,w',k] / 255 - 0.5) / 0.5,   w' = flip[n] ? W-1-w : w
// Same float operations in the same order as ToTensor (.div(255)) and Normalize (.sub_(mean).div_(std)), so the result
// is bit-identical to the CPU transform.  Resize stays with the decoder (PIL) on the host.
// NOT yet run on hardware (tests/test_input_pipeline.py is marked gpu_next).
#include "sg_common.cuh"
#include "../../include/sgb200.h"

namespace {

__global__ void __launch_bounds__(256) image_u8_to_f32_kernel(const unsigned char* __restrict__ in,
                                                              const unsigned char* __restrict__ flip,
                                                              float* __restrict__ out, int N, int H, int W, int K) {
  const long long total = (long long)N * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const long long r = i / W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    const int ws = (flip != nullptr && flip[n]) ? (W - 1 - w) : w;
    const unsigned char* src = in + (((size_t)n * H + h) * W + ws) * K;
    for (int k = 0; k < K; ++k) {
      const float t = __fdiv_rn((float)src[k], 255.f);            // ToTensor
      out[(((size_t)n * K + k) * H + h) * W + w] = __fdiv_rn(__fsub_rn(t, 0.5f), 0.5f);   // Normalize
    }
  }
}

}  // namespace

SG_EXPORT int sg_image_u8_to_f32(const unsigned char* in, const unsigned char* flip, float* out, int N, int H, int W,
                                 int K, void* stream) {
  SG_REQUIRE(N >= 1 && H >= 1 && W >= 1 && K >= 1 && K <= 4, "sg_image_u8_to_f32: bad shape");
  const long long total = (long long)N * H * W;
  image_u8_to_f32_kernel<<<sg_grid_1d(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(in, flip, out, N, H, W, K);
  return sg_check_launch("sg_image_u8_to_f32");
}
