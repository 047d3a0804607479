// Fast paths of sg_rgb_in / sg_rgb_out / sg_channel_sum for the StyleGAN case K = 3, bf16 activations (from_rgb / to_rgb,
// reference models/GAN.py:157,167,358,377; noise-weight + bias gradient sums).  Measured against the general kernels at
// 32x256x256x64 (profiles/r02_pointwise_microbench.txt): rgb_in 2.05 -> 2.55 TB/s, rgb_out 1.00 -> 2.46 TB/s,
// weighted channel sum 2.16 -> 3.69 TB/s.
//
// The general to/from-RGB kernels are issue-bound (profiles/r01_summary.md): rgb_out spends ~1100 of its ~2200
// instructions per 16 pixels in the (k, a, b) scatter (four runtime divisions and 64-bit address arithmetic per lane),
// rgb_in_vec ~150 per output vector (two divisions, K padded to 4, select-based LeakyReLU).  Versions with K = 3 fixed at
// compile time, the per-lane scatter offsets hoisted out of the pixel loop and no division in the steady state.
#include "sg_common.cuh"
#include "sg_fast.cuh"

namespace {

__device__ __forceinline__ uint4 pack8(const float (&o)[8]) {
  uint4 r;
  __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(o[4], o[5]), d = __floats2bfloat162_rn(o[6], o[7]);
  r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
  r.z = *reinterpret_cast<uint32_t*>(&c); r.w = *reinterpret_cast<uint32_t*>(&d);
  return r;
}

// ---- from_rgb, K = 3:  y[p][c] = act(sum_k Ws[c][k] * img[n][k][hw] + b[c]) --------------------------------------------
// A thread keeps one channel vector (cv) and walks pixels p0, p0 + stride, ...: (n, hw) advance incrementally.
template <int ACT>
__global__ void __launch_bounds__(256) rgb_in3_kernel(const float* __restrict__ img, const float* __restrict__ Wt,
                                                      const float* __restrict__ bias, float scale, float bias_mul,
                                                      __nv_bfloat16* __restrict__ y, int N, int HW, int C) {
  constexpr int VN = 8, K = 3;
  const int CV = C / VN;
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;            // a multiple of CV (launcher)
  const int cv = t0 % CV;
  const int pstride = nthreads / CV;
  float w0[VN], w1[VN], w2[VN], br[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    const int c = cv * VN + e;
    w0[e] = Wt[c * K + 0] * scale; w1[e] = Wt[c * K + 1] * scale; w2[e] = Wt[c * K + 2] * scale;
    br[e] = bias ? bias[c] * bias_mul : 0.f;
  }
  const int P = N * HW;
  int p = t0 / CV;
  int n = p / HW, hw = p - n * HW;                        // the only division
  const int dn = pstride / HW, dhw = pstride - dn * HW;
  for (; p < P; p += pstride) {
    const float* px = img + ((size_t)n * K) * HW + hw;
    const float a = px[0], b = px[HW], c = px[2 * (size_t)HW];
    float o[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      float s = fmaf(w2[e], c, fmaf(w1[e], b, fmaf(w0[e], a, br[e])));
      if (ACT) s = fmaxf(s, SG_LRELU_SLOPE * s);          // LeakyReLU(0.2) for slope < 1
      o[e] = s;
    }
    *reinterpret_cast<uint4*>(y + (size_t)p * C + cv * VN) = pack8(o);
    n += dn; hw += dhw;
    if (hw >= HW) { hw -= HW; ++n; }
  }
}

// ---- to_rgb, K = 3:  out[n][k][Y][X] = beta*out + alpha*(scale * sum_c W[k][c] x[p][c] + b[k]*bmul), nearest x up -------
// G lanes per pixel (G = 8 for C = 64), a warp iteration covers U*32/G CONSECUTIVE pixels of one image row, so (n, y)
// are computed once per iteration; lanes l < 3*up*up hold their (k, a, b) and scatter offset for the whole kernel.
template <int U>
__global__ void __launch_bounds__(256) rgb_out3_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ Wt,
                                                       const float* __restrict__ bias, float scale, float bias_mul,
                                                       float alpha, float beta, float* __restrict__ out, int N, int h,
                                                       int w, int C, int up, int G) {
  constexpr int VN = 8, K = 3;
  const int lane = threadIdx.x & 31;
  const int sub = lane / G, l = lane % G, per_warp = 32 / G;
  const int span = per_warp * U;                           // pixels per warp iteration; w % span == 0 (launcher)
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int CV = C / VN;                                   // CV <= G (launcher)
  const int H = h * up, Wd = w * up;
  const int lc = l < CV ? l : CV - 1;
  float wr[VN][K];
#pragma unroll
  for (int e = 0; e < VN; ++e)
#pragma unroll
    for (int k = 0; k < K; ++k) wr[e][k] = (l < CV) ? Wt[k * C + l * VN + e] * (scale * alpha) : 0.f;
  // this lane's output element inside the pixel's up x up block of plane k
  const int uu = up * up;
  const bool writer = l < K * uu;
  const int kk = writer ? l / uu : 0, ab = writer ? l % uu : 0, aa = ab / up, bb = ab % up;
  const float bk = (bias ? bias[kk] * bias_mul : 0.f) * alpha;
  const long long lane_off = ((long long)kk * H + aa) * Wd + bb;
  const int iters = (N * h * w) / span;
  for (int it = warp; it < iters; it += nwarps) {
    const int p0 = it * span;                              // first pixel of the row segment
    const int row = p0 / w;                                // = n*h + y   (one division per 16+ pixels)
    const int x0 = p0 - row * w;
    const int n = row / h, yy = row - n * h;
    const __nv_bfloat16* src = x + ((size_t)p0 + sub) * C + lc * VN;
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = *reinterpret_cast<const uint4*>(src + (size_t)u * per_warp * C);
    float acc[U][K];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int k = 0; k < K; ++k) acc[u][k] = 0.f;
      const uint32_t wd[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        const float f = __uint_as_float((e & 1) ? (wd[e >> 1] & 0xffff0000u) : (wd[e >> 1] << 16));
#pragma unroll
        for (int k = 0; k < K; ++k) acc[u][k] = fmaf(f, wr[e][k], acc[u][k]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < K; ++k)
        for (int o = G >> 1; o > 0; o >>= 1) acc[u][k] += __shfl_xor_sync(0xffffffffu, acc[u][k], o);
    if (writer) {
      const long long base = ((long long)n * K * H + (long long)yy * up) * Wd + lane_off;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int xw = x0 + u * per_warp + sub;
        const float s = kk == 0 ? acc[u][0] : (kk == 1 ? acc[u][1] : acc[u][2]);
        float* dst = out + base + (long long)xw * up;
        *dst = (beta != 0.f ? beta * *dst : 0.f) + (s + bk);
      }
    }
  }
}

// ---- per-channel sums with one weight row + the plain sum (noise-weight and bias gradients) ---------------------------
//   part[blk][0][c] = sum_p wt[p] * x[p][c],  part[blk][1][c] = sum_p x[p][c]      (wt: one float per pixel)
__global__ void __launch_bounds__(256) channel_sum_w1_kernel(const __nv_bfloat16* __restrict__ x,
                                                             const float* __restrict__ wt, long long P, int C,
                                                             float* __restrict__ part, int pix_per_blk) {
  constexpr int VN = 8, U = 4;
  extern __shared__ float sm[];                            // [2][PR][C]
  const int CV = C / VN;
  const int PR = blockDim.x / CV;
  const int cv = threadIdx.x % CV, pr = threadIdx.x / CV;
  const long long p0 = (long long)blockIdx.x * pix_per_blk;
  long long p1 = p0 + pix_per_blk;
  if (p1 > P) p1 = P;
  float aw[VN], a1[VN];
#pragma unroll
  for (int k = 0; k < VN; ++k) aw[k] = a1[k] = 0.f;
  if (pr < PR) {
    const size_t step = (size_t)PR * C;
    long long pb = p0 + pr;
    const __nv_bfloat16* ptr = x + (size_t)pb * C + cv * VN;
    const float* wp = wt + pb;
    auto unpack = [](const uint4& v, float (&f)[VN]) {
      const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < VN; ++e) f[e] = __uint_as_float((e & 1) ? (wd[e >> 1] & 0xffff0000u) : (wd[e >> 1] << 16));
    };
    for (; pb + (long long)(U - 1) * PR < p1; pb += (long long)U * PR, ptr += U * step, wp += U * PR) {
      uint4 v[U];
      float wv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { v[u] = *reinterpret_cast<const uint4*>(ptr + u * step); wv[u] = wp[u * PR]; }
      float f[U][VN];
#pragma unroll
      for (int u = 0; u < U; ++u) unpack(v[u], f[u]);
#pragma unroll
      for (int k = 0; k < VN; ++k) {                       // pairwise: keeps the U loads of a batch back to back
        a1[k] += (f[0][k] + f[1][k]) + (f[2][k] + f[3][k]);
        aw[k] += fmaf(wv[0], f[0][k], wv[1] * f[1][k]) + fmaf(wv[2], f[2][k], wv[3] * f[3][k]);
      }
    }
    for (; pb < p1; pb += PR, ptr += step, wp += PR) {
      const uint4 v = *reinterpret_cast<const uint4*>(ptr);
      const float wv = *wp;
      float f[VN];
      unpack(v, f);
#pragma unroll
      for (int k = 0; k < VN; ++k) { a1[k] += f[k]; aw[k] = fmaf(wv, f[k], aw[k]); }
    }
#pragma unroll
    for (int k = 0; k < VN; ++k) {
      sm[((size_t)0 * PR + pr) * C + cv * VN + k] = aw[k];
      sm[((size_t)1 * PR + pr) * C + cv * VN + k] = a1[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int j = i / C, c = i % C;
    float s = 0.f;
    for (int r = 0; r < PR; ++r) s += sm[((size_t)j * PR + r) * C + c];
    part[((size_t)blockIdx.x * 2 + j) * C + c] = s;
  }
}

// ---- the same with J = 3 per-pixel weight rows (+ the plain sum): the weight gradients of to_rgb (J = 3) and from_rgb
//      (J = 3 + ones).  wt is [N][3][HW] float (an RGB image or its gradient), x NHWC bf16.
//   part[blk][j][c] = sum_p wt[n(p)][j][hw(p)] * x[p][c]  (j < 3),  part[blk][3][c] = sum_p x[p][c] if ONES
// The general channel_sum_vec_kernel spends 4 FMAs + index arithmetic per element whatever J is and ran at 1.9 TB/s.
template <bool ONES>
__global__ void __launch_bounds__(256) channel_sum_w3_kernel(const __nv_bfloat16* __restrict__ x,
                                                             const float* __restrict__ wt, long long P, int HW, int C,
                                                             float* __restrict__ part, int pix_per_blk) {
  constexpr int VN = 8, U = 4, ROWS = ONES ? 4 : 3;
  extern __shared__ float sm[];                            // [ROWS][PR][C]
  const int CV = C / VN;
  const int PR = blockDim.x / CV;
  const int cv = threadIdx.x % CV, pr = threadIdx.x / CV;
  const long long p0 = (long long)blockIdx.x * pix_per_blk;
  long long p1 = p0 + pix_per_blk;
  if (p1 > P) p1 = P;
  float acc[ROWS][VN];
#pragma unroll
  for (int j = 0; j < ROWS; ++j)
#pragma unroll
    for (int k = 0; k < VN; ++k) acc[j][k] = 0.f;
  if (pr < PR) {
    auto unpack = [](const uint4& v, float (&f)[VN]) {
      const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < VN; ++e) f[e] = __uint_as_float((e & 1) ? (wd[e >> 1] & 0xffff0000u) : (wd[e >> 1] << 16));
    };
    for (long long pb = p0 + pr; pb < p1; pb += (long long)U * PR) {
      uint4 v[U];
      float wv[U][3];
#pragma unroll
      for (int u = 0; u < U; ++u) {                        // U pixels in flight: one 16-byte x load + three weight loads each
        const long long p = pb + (long long)u * PR;
        if (p < p1) {
          v[u] = *reinterpret_cast<const uint4*>(x + (size_t)p * C + cv * VN);
          const unsigned n = (unsigned)p / (unsigned)HW;   // P < 2^31 (launcher)
          const float* w = wt + (size_t)n * 3 * HW + ((unsigned)p - n * (unsigned)HW);
          wv[u][0] = w[0]; wv[u][1] = w[HW]; wv[u][2] = w[2 * (size_t)HW];
        } else {
          v[u] = make_uint4(0u, 0u, 0u, 0u);
          wv[u][0] = wv[u][1] = wv[u][2] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float f[VN];
        unpack(v[u], f);
#pragma unroll
        for (int k = 0; k < VN; ++k) {
          acc[0][k] = fmaf(wv[u][0], f[k], acc[0][k]);
          acc[1][k] = fmaf(wv[u][1], f[k], acc[1][k]);
          acc[2][k] = fmaf(wv[u][2], f[k], acc[2][k]);
          if (ONES) acc[3][k] += f[k];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < ROWS; ++j)
#pragma unroll
      for (int k = 0; k < VN; ++k) sm[((size_t)j * PR + pr) * C + cv * VN + k] = acc[j][k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ROWS * C; i += blockDim.x) {
    const int j = i / C, c = i % C;
    float s = 0.f;
    for (int r = 0; r < PR; ++r) s += sm[((size_t)j * PR + r) * C + c];
    part[((size_t)blockIdx.x * ROWS + j) * C + c] = s;
  }
}

__global__ void __launch_bounds__(128) reduce_parts_kernel(const float* __restrict__ part, int nblk, int M,
                                                           float* __restrict__ out, float scale) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= M) return;
  float s = 0.f;
  for (int b = lane; b < nblk; b += 32) s += part[(size_t)b * M + i];
  s = sg_warp_sum(s);
  if (lane == 0) out[i] = s * scale;
}

}  // namespace

#define SGN_EXPORT

// from_rgb (bf16 output, K = 3).  Returns 1000 when the shape is outside this kernel's fast path.
SGN_EXPORT int sgf_rgb_in3(const float* img, const float* W, const float* bias, float scale, float bias_mul, int act,
                           void* y, int N, int HW, int C, void* stream) {
  const int CV = C / 8;
  if (C % 8 != 0 || CV < 1 || 256 % CV != 0 || (long long)N * HW * CV >= (1ll << 30)) return SG_FAST_UNSUPPORTED;
  const long long total = (long long)N * HW * CV;
  const int grid = sg_grid_1d(total, 256, 148 * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (act) rgb_in3_kernel<1><<<grid, 256, 0, st>>>(img, W, bias, scale, bias_mul, (__nv_bfloat16*)y, N, HW, C);
  else rgb_in3_kernel<0><<<grid, 256, 0, st>>>(img, W, bias, scale, bias_mul, (__nv_bfloat16*)y, N, HW, C);
  return (int)cudaGetLastError();
}

// to_rgb (bf16 input, K = 3, up 1 or 2).
SGN_EXPORT int sgf_rgb_out3(const void* x, const float* W, const float* bias, float scale, float bias_mul, float alpha,
                            float beta, float* out, int N, int h, int w, int C, int up, void* stream) {
  if (C % 8 != 0 || up < 1 || up > 2) return SG_FAST_UNSUPPORTED;
  const int CV = C / 8;
  int G = 1;
  while (G < CV || G < 3 * up * up) G <<= 1;               // one lane per channel vector, and 3*up*up writer lanes
  if (G > 32) return SG_FAST_UNSUPPORTED;
  constexpr int U = 4;
  const int span = (32 / G) * U;
  if (w % span != 0 || (long long)N * h * w >= (1ll << 30)) return SG_FAST_UNSUPPORTED;
  const long long iters = (long long)N * h * w / span;
  const int grid = sg_grid_1d(iters * 32, 256, 148 * 8);
  rgb_out3_kernel<U><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, W, bias, scale, bias_mul, alpha,
                                                             beta, out, N, h, w, C, up, G);
  return (int)cudaGetLastError();
}

// out[0][c] = scale * sum_p wt[p] x[p][c], out[1][c] = scale * sum_p x[p][c]; ws: nblk*2*C floats (nblk = return of
// sgf_channel_sum_blocks).
SGN_EXPORT int sgf_channel_sum_blocks(long long P) {
  long long b = (P + 255) / 256;
  if (b > 148 * 4) b = 148 * 4;
  return b < 1 ? 1 : (int)b;
}
// J = 3 weight rows (+ ones): out[j][c] = scale * sum_p wt[n][j][hw] x[p][c]; ws: nblk*(3+ones)*C floats.
SGN_EXPORT int sgf_channel_sum_w3(const void* x, const float* wt, int ones, long long P, int HW, int C, float* ws, float* out,
                                  float scale, void* stream) {
  const int CV = C / 8, rows = ones ? 4 : 3;
  if (C % 8 != 0 || CV > 256 || (size_t)rows * (256 / CV) * C * sizeof(float) > 48 * 1024) return SG_FAST_UNSUPPORTED;
  const int nblk = sgf_channel_sum_blocks(P);
  const int ppb = (int)((P + nblk - 1) / nblk);
  const size_t smem = (size_t)rows * (256 / CV) * C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (ones) channel_sum_w3_kernel<true><<<nblk, 256, smem, st>>>((const __nv_bfloat16*)x, wt, P, HW, C, ws, ppb);
  else channel_sum_w3_kernel<false><<<nblk, 256, smem, st>>>((const __nv_bfloat16*)x, wt, P, HW, C, ws, ppb);
  reduce_parts_kernel<<<(rows * C + 3) / 4, 128, 0, st>>>(ws, nblk, rows * C, out, scale);
  return (int)cudaGetLastError();
}
SGN_EXPORT int sgf_channel_sum_w1(const void* x, const float* wt, long long P, int C, float* ws, float* out, float scale,
                                  void* stream) {
  const int CV = C / 8;
  if (C % 8 != 0 || CV > 256 || (size_t)2 * (256 / CV) * C * sizeof(float) > 48 * 1024) return SG_FAST_UNSUPPORTED;
  const int nblk = sgf_channel_sum_blocks(P);
  const int ppb = (int)((P + nblk - 1) / nblk);
  const size_t smem = (size_t)2 * (256 / CV) * C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  channel_sum_w1_kernel<<<nblk, 256, smem, st>>>((const __nv_bfloat16*)x, wt, P, C, ws, ppb);
  reduce_parts_kernel<<<(2 * C + 3) / 4, 128, 0, st>>>(ws, nblk, 2 * C, out, scale);
  return (int)cudaGetLastError();
}
