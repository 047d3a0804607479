// Fast path of sg_blur3 (BlurLayer, reference models/CustomLayers.py:266-276): measured 4.93 TB/s against 3.08 TB/s
// for the general strip kernel at 32x256x256x64 bf16 (profiles/r02_pointwise_microbench.txt).
//
// blur3 v2: the general kernel (csrc/sg_pointwise.cu) is issue-bound at ~206 instructions per 16-byte output vector.  Here a thread owns FOUR adjacent pixels of one 16-byte channel vector and an 8-row strip: per input row
// it unpacks 6 neighbour vectors for 4 outputs (1.5 instead of 3 per output), border columns are zeroed on the packed
// words of the two edge vectors only, and only the first / last row of a strip can be outside the image.
// Static count: ~75 instructions per output vector.  Requires W % 4 == 0, H % 8 == 0, C % (16 / sizeof(T)) == 0.
#include "sg_common.cuh"
#include "sg_fast.cuh"

namespace {

constexpr int STRIP = 8;      // output rows per thread
constexpr int PX = 4;         // adjacent pixels per thread
constexpr int ROWS_PER_BATCH = 2;

template <class V> __device__ __forceinline__ void zero_if(V& v, bool z) {
  if (z) v.zero();
}

// one cvt.rn.bf16x2.f32 per pair instead of a convert + byte-permute per element
__device__ __forceinline__ void store_vec(__nv_bfloat16* p, const float (&o)[8]) {
  uint4 r;
  __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(o[4], o[5]), d = __floats2bfloat162_rn(o[6], o[7]);
  r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
  r.z = *reinterpret_cast<uint32_t*>(&c); r.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(p) = r;
}
__device__ __forceinline__ void store_vec(float* p, const float (&o)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}

template <class T>
__global__ void __launch_bounds__(128, 3) blur3_v2_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                       const float* __restrict__ bias, float bias_mul, int N, int H,
                                                       int W, int C) {
  using V = SgVec<T>;
  constexpr int VN = V::N;
  const int CV = C / VN, WG = W / PX, strips = H / STRIP;
  const int total = N * strips * WG * CV;              // < 2^31 checked by the launcher
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cv = i % CV;
    int r = i / CV;
    const int g = r % WG; r /= WG;
    const int st = r % strips;
    const int n = r / strips;
    const int w0 = g * PX, h0 = st * STRIP;
    const bool edge_l = (w0 == 0), edge_r = (w0 + PX == W);
    float bb[VN];
#pragma unroll
    for (int k = 0; k < VN; ++k) bb[k] = bias ? bias[cv * VN + k] * bias_mul : 0.f;
    // column j of the 6-wide window is pixel w0 - 1 + j; the two edge columns are clamped and zeroed when outside
    const T* base = x + ((size_t)n * H * W) * C + cv * VN;
    int col_off[PX + 2];
#pragma unroll
    for (int j = 0; j < PX + 2; ++j) {
      int w = w0 - 1 + j;
      w = w < 0 ? 0 : (w >= W ? W - 1 : w);
      col_off[j] = w * C;
    }
    float r0[PX][VN] = {}, r1[PX][VN] = {};             // horizontal sums of the two rows above
#pragma unroll
    for (int b = 0; b < (STRIP + 2) / ROWS_PER_BATCH; ++b) {
      V raw[ROWS_PER_BATCH][PX + 2];
#pragma unroll
      for (int q = 0; q < ROWS_PER_BATCH; ++q) {
        int hh = h0 - 1 + b * ROWS_PER_BATCH + q;
        hh = hh < 0 ? 0 : (hh >= H ? H - 1 : hh);
        const T* row = base + (size_t)hh * W * C;
#pragma unroll
        for (int j = 0; j < PX + 2; ++j) raw[q][j].load(row + col_off[j]);
      }
#pragma unroll
      for (int q = 0; q < ROWS_PER_BATCH; ++q) {
        const int ri = b * ROWS_PER_BATCH + q;          // input row h0 - 1 + ri  (compile-time)
        zero_if(raw[q][0], edge_l);
        zero_if(raw[q][PX + 1], edge_r);
        if (ri == 0 || ri == STRIP + 1) {               // only these rows of a strip can lie outside the image
          const int hh = h0 - 1 + ri;
          const bool out_of_image = hh < 0 || hh >= H;
#pragma unroll
          for (int j = 0; j < PX + 2; ++j) zero_if(raw[q][j], out_of_image);
        }
        float f[PX + 2][VN];
#pragma unroll
        for (int j = 0; j < PX + 2; ++j)
#pragma unroll
          for (int k = 0; k < VN; ++k) f[j][k] = raw[q][j].get(k);
        float hs[PX][VN];
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
          for (int k = 0; k < VN; ++k) hs[p][k] = (2.f * f[p + 1][k] + f[p][k]) + f[p + 2][k];
        if (ri >= 2) {
          const int h = h0 + ri - 2;                    // always < H because H % STRIP == 0
          T* dst = y + (((size_t)n * H + h) * W + w0) * C + cv * VN;
#pragma unroll
          for (int p = 0; p < PX; ++p) {
            float o[VN];
#pragma unroll
            for (int k = 0; k < VN; ++k) o[k] = (r0[p][k] + 2.f * r1[p][k] + hs[p][k]) * (1.f / 16.f) + bb[k];
            store_vec(dst + (size_t)p * C, o);
          }
        }
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
          for (int k = 0; k < VN; ++k) { r0[p][k] = r1[p][k]; r1[p][k] = hs[p][k]; }
      }
    }
  }
}

}  // namespace

// -> 0 launched, SG_FAST_UNSUPPORTED when the shape is outside this kernel's domain (caller falls back), else cudaError
int sgf_blur3(const void* x, void* y, const float* bias, float bias_mul, int dtype, int N, int H, int W, int C,
              void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (H % STRIP != 0 || W % PX != 0) return SG_FAST_UNSUPPORTED;
  if (dtype == SG_BF16) {
    using T = __nv_bfloat16;
    if (C % SgVec<T>::N != 0) return SG_FAST_UNSUPPORTED;
    long long total = (long long)N * (H / STRIP) * (W / PX) * (C / SgVec<T>::N);
    if (total >= (1ll << 30)) return SG_FAST_UNSUPPORTED;
    blur3_v2_kernel<T><<<sg_grid_1d(total, 128, 148 * 16), 128, 0, st>>>((const T*)x, (T*)y, bias, bias_mul, N, H, W, C);
  } else if (dtype == SG_F32) {
    using T = float;
    if (C % SgVec<T>::N != 0) return SG_FAST_UNSUPPORTED;
    long long total = (long long)N * (H / STRIP) * (W / PX) * (C / SgVec<T>::N);
    if (total >= (1ll << 30)) return SG_FAST_UNSUPPORTED;
    blur3_v2_kernel<T><<<sg_grid_1d(total, 128, 148 * 16), 128, 0, st>>>((const T*)x, (T*)y, bias, bias_mul, N, H, W, C);
  } else {
    return SG_FAST_UNSUPPORTED;
  }
  return (int)cudaGetLastError();
}
