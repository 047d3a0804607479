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


// ---- blur + the front half of LayerEpilogue (reference CustomLayers.py:174-180 then :242-246) in ONE pass -------------
//   z = lrelu( blur3(x) + bias[c]*bias_mul + noise_w[c]*noise[n,h,w] )  (bf16)
//   part[n][chunk][0][c] = sum of z over the chunk's pixels, part[n][chunk][1][c] = sum of z^2   (of the ROUNDED z)
// grid (chunks, N): a block stays inside one image and a thread keeps one channel vector, so the statistics are private
// register sums reduced once per block through shared memory (fixed order: deterministic, no atomics).
template <int DUMMY>
__global__ void __launch_bounds__(128, 2) blur3_epi_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ z,
                                                        const float* __restrict__ bias, float bias_mul,
                                                        const float* __restrict__ noise, const float* __restrict__ noise_w,
                                                        float* __restrict__ part, int H, int W, int C,
                                                        int items_per_blk) {
  using T = __nv_bfloat16;
  using V = SgVec<T>;
  constexpr int VN = V::N;
  extern __shared__ float sm[];                        // [2][PR][C]
  const int CV = C / VN, WG = W / PX, strips = H / STRIP;
  const int items = strips * WG * CV;                  // per image
  const int n = blockIdx.y;
  const int cv = threadIdx.x % CV, pr = threadIdx.x / CV, PR = blockDim.x / CV;   // blockDim.x % CV == 0 (launcher)
  const int i_end = min(items, (int)(blockIdx.x + 1) * items_per_blk);
  float bb[VN], nw[VN], s1[VN], s2[VN];
#pragma unroll
  for (int k = 0; k < VN; ++k) {
    bb[k] = bias ? bias[cv * VN + k] * bias_mul : 0.f;
    nw[k] = noise ? noise_w[cv * VN + k] : 0.f;
    s1[k] = s2[k] = 0.f;
  }
  const T* base = x + ((size_t)n * H * W) * C + cv * VN;
  for (int i = blockIdx.x * items_per_blk + threadIdx.x; i < i_end; i += blockDim.x) {   // i % CV == cv
    int r = i / CV;
    const int g = r % WG;
    const int st = r / WG;
    const int w0 = g * PX, h0 = st * STRIP;
    const bool edge_l = (w0 == 0), edge_r = (w0 + PX == W);
    int col_off[PX + 2];
#pragma unroll
    for (int j = 0; j < PX + 2; ++j) {
      int w = w0 - 1 + j;
      w = w < 0 ? 0 : (w >= W ? W - 1 : w);
      col_off[j] = w * C;
    }
    float r0[PX][VN] = {}, r1[PX][VN] = {};
#pragma unroll
    for (int b = 0; b < (STRIP + 2) / ROWS_PER_BATCH; ++b) {
      V raw[ROWS_PER_BATCH][PX + 2];
      float4 nzv[ROWS_PER_BATCH];
#pragma unroll
      for (int q = 0; q < ROWS_PER_BATCH; ++q) {
        int hh = h0 - 1 + b * ROWS_PER_BATCH + q;
        hh = hh < 0 ? 0 : (hh >= H ? H - 1 : hh);
        const T* row = base + (size_t)hh * W * C;
#pragma unroll
        for (int j = 0; j < PX + 2; ++j) raw[q][j].load(row + col_off[j]);
        const int ri = b * ROWS_PER_BATCH + q;
        nzv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (noise && ri >= 2)                            // noise of OUTPUT row h0 + ri - 2, pixels w0..w0+3
          nzv[q] = *reinterpret_cast<const float4*>(noise + ((size_t)n * H + (h0 + ri - 2)) * W + w0);
      }
#pragma unroll
      for (int q = 0; q < ROWS_PER_BATCH; ++q) {
        const int ri = b * ROWS_PER_BATCH + q;
        zero_if(raw[q][0], edge_l);
        zero_if(raw[q][PX + 1], edge_r);
        if (ri == 0 || ri == STRIP + 1) {
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
          const int h = h0 + ri - 2;
          T* dst = z + (((size_t)n * H + h) * W + w0) * C + cv * VN;
          const float nzp[PX] = {nzv[q].x, nzv[q].y, nzv[q].z, nzv[q].w};
#pragma unroll
          for (int p = 0; p < PX; ++p) {
            float o[VN];
#pragma unroll
            for (int k = 0; k < VN; ++k) {
              float v = fmaf(nw[k], nzp[p], (r0[p][k] + 2.f * r1[p][k] + hs[p][k]) * (1.f / 16.f) + bb[k]);
              o[k] = fmaxf(v, SG_LRELU_SLOPE * v);
            }
            V ov;
#pragma unroll
            for (int k = 0; k < VN; ++k) ov.set(k, o[k]);
            ov.store(dst + (size_t)p * C);
#pragma unroll
            for (int k = 0; k < VN; ++k) {
              const float zr = ov.get(k);
              s1[k] += zr;
              s2[k] = fmaf(zr, zr, s2[k]);
            }
          }
        }
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
          for (int k = 0; k < VN; ++k) { r0[p][k] = r1[p][k]; r1[p][k] = hs[p][k]; }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < VN; ++k) {
    sm[((size_t)0 * PR + pr) * C + cv * VN + k] = s1[k];
    sm[((size_t)1 * PR + pr) * C + cv * VN + k] = s2[k];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * C; t += blockDim.x) {
    const int jq = t / C, c = t % C;
    float s = 0.f;
    for (int rr = 0; rr < PR; ++rr) s += sm[((size_t)jq * PR + rr) * C + c];
    part[(((size_t)n * gridDim.x + blockIdx.x) * 2 + jq) * C + c] = s;
  }
}

// all slots valid: per (n, c) sum over `slots` partial records in fixed order (fp64), then the folded affine
__global__ void finalize_slots_kernel(const float* __restrict__ part, int slots, const float* __restrict__ style,
                                      float* __restrict__ mr, float* __restrict__ ab, int N, int C, int HW, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i % C;
  double a = 0.0, b = 0.0;
  for (int s = 0; s < slots; ++s) {
    const float* q = part + (((size_t)n * slots + s) * 2) * C + c;
    a += (double)q[0];
    b += (double)q[C];
  }
  const double m = a / HW;
  double var = b / HW - m * m;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  mr[(size_t)i * 2] = (float)m;
  mr[(size_t)i * 2 + 1] = rstd;
  const float st0 = style ? style[(size_t)n * 2 * C + c] : 0.f;
  const float st1 = style ? style[(size_t)n * 2 * C + C + c] : 0.f;
  ab[(size_t)i * 3] = (float)m;
  ab[(size_t)i * 3 + 1] = rstd * (1.f + st0);
  ab[(size_t)i * 3 + 2] = st1;
}

static bool blur_epi_plan(int N, int H, int W, int C, int* chunks, int* items_per_blk, int* threads) {
  if (H % STRIP != 0 || W % PX != 0 || C % 8 != 0 || N > 65535) return false;
  const int CV = C / 8;
  if (CV > 128 || 128 % CV != 0) return false;
  if ((size_t)2 * (128 / CV) * C * sizeof(float) > 48 * 1024) return false;
  const int items = (H / STRIP) * (W / PX) * CV;
  int c = (148 * 6 + N - 1) / N;                       // ~6 blocks per SM over the whole launch
  int ipb = (items + c - 1) / c;
  ipb = ((ipb + 127) / 128) * 128;                     // keeps a thread on one channel vector
  *items_per_blk = ipb;
  *chunks = (items + ipb - 1) / ipb;
  *threads = 128;
  return true;
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

#include "../../include/sgb200.h"
SG_EXPORT int sg_blur3_epi_slots(int N, int H, int W, int C) {
  int chunks, ipb, thr;
  return blur_epi_plan(N, H, W, C, &chunks, &ipb, &thr) ? chunks : 0;
}
// bf16 NHWC; noise fp32 [N,H,W] or null (then noise_w is ignored); part fp32 [N][sg_blur3_epi_slots][2][C]
SG_EXPORT int sg_blur3_epi(const void* x, void* z, const float* bias, float bias_mul, const float* noise,
                           const float* noise_w, float* part, int N, int H, int W, int C, void* stream) {
  int chunks, ipb, thr;
  SG_REQUIRE(blur_epi_plan(N, H, W, C, &chunks, &ipb, &thr), "sg_blur3_epi: shape not supported");
  SG_REQUIRE(part != nullptr, "sg_blur3_epi: part buffer missing");
  const size_t smem = (size_t)2 * (128 / (C / 8)) * C * sizeof(float);
  blur3_epi_kernel<0><<<dim3(chunks, N), thr, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (__nv_bfloat16*)z, bias, bias_mul, noise, noise_w, part, H, W, C, ipb);
  return sg_check_launch("sg_blur3_epi");
}
SG_EXPORT int sg_epi_finalize_slots(const float* part, int slots, const float* style, float* mr, float* ab, int N, int C,
                                    int HW, float eps, void* stream) {
  finalize_slots_kernel<<<sg_cdiv((long long)N * C, 128), 128, 0, (cudaStream_t)stream>>>(part, slots, style, mr, ab, N,
                                                                                          C, HW, eps);
  return sg_check_launch("sg_epi_finalize_slots");
}
