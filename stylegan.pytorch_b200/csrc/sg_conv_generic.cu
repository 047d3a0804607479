// Generic tap-table convolution family (CUDA-core FFMA, fp32 accumulate) on NHWC activations.
//
// One geometry struct describes every dense convolution on the StyleGAN hot path and its adjoints:
//   out[n, qy*os+py, qx*os+px, co] = sum_t sum_ci in[n, qy*is+dy[ph][t], qx*is+dx[ph][t], ci] * E[ph][t][ci][co]
//   (is,os)=(1,1): 3x3/1x1 stride-1 'same' conv (and its dgrad)            reference CustomLayers.py:170-173
//   (is,os)=(2,1): 4x4 stride-2 pad-1 conv = conv3x3 -> 2x2 box            reference CustomLayers.py:156-165, 48-64
//   (is,os)=(1,2): 4 output phases x 2x2 taps = nearest-up x2 -> conv3x3   reference CustomLayers.py:143-154
//                  (== conv_transpose2d of the box-summed 4x4 kernel)
// The effective weights E are produced from the 3x3 parameter by sg_wprep_fwd with a coefficient table built on
// the host (models/_geom.py); the same table transposed maps dE back to dW (sg_wprep_bwd).
//
// This is the fp32 parity path and the universal fallback; the bf16 hot layers run sg_conv_tc.cu (tcgen05).
#include "sg_common.cuh"
#include <mutex>
#include "../../include/sgb200.h"

namespace {

constexpr int BM = 64;   // pixels (or ci for wgrad) per CTA tile
constexpr int BN = 64;   // output channels per CTA tile
constexpr int BK = 16;   // reduction chunk
constexpr int NT = 256;  // threads
constexpr int APAD = 4;

template <class T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  float4 r = *reinterpret_cast<const float4*>(p);
  v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
  v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
}
template <class T> __device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// 4 consecutive channels starting at c (c may run past C: tail handled element-wise when !VEC)
template <class T, bool VEC>
__device__ __forceinline__ void load_ch4(const T* base, int c, int C, bool valid, float (&v)[4]) {
  if (!valid) { v[0] = v[1] = v[2] = v[3] = 0.f; return; }
  if (VEC) {
    if (c < C) load4<T>(base + c, v); else { v[0] = v[1] = v[2] = v[3] = 0.f; }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (c + i < C) ? sg_to_f(base[c + i]) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------------------
// forward / dgrad: implicit GEMM  [pixels x Cin*taps] x [Cin*taps x Cout]
// ------------------------------------------------------------------------------------------------------------
template <class T, bool VEC>
__global__ void __launch_bounds__(NT) conv_generic_kernel(const T* __restrict__ x, const T* __restrict__ E,
                                                          T* __restrict__ y, const float* __restrict__ bias,
                                                          float bias_mul, int act, SgConvGeom g) {
  __shared__ float As[BK][BM + APAD];
  __shared__ float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int ph = blockIdx.z;
  const int py = ph >> 1, px = ph & 1;
  const long long Q = (long long)g.N * g.Hq * g.Wq;
  const long long m0 = (long long)blockIdx.x * BM;
  const int co0 = blockIdx.y * BN;

  // A-load role: pixel a_m (0..63), channel chunk a_c (0,4,8,12)
  const int a_m = tid >> 2, a_c = (tid & 3) * 4;
  long long am = m0 + a_m;
  bool a_ok = am < Q;
  int an = 0, aqy = 0, aqx = 0;
  if (a_ok) {
    aqx = (int)(am % g.Wq);
    long long r = am / g.Wq;
    aqy = (int)(r % g.Hq);
    an = (int)(r / g.Hq);
  }
  // B-load role: ci row b_k (0..15), co chunk b_c
  const int b_k = tid >> 4, b_c = (tid & 15) * 4;
  // compute role: 4 pixels x 4 couts
  const int ty = tid >> 4, tx = tid & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const T* Eph = E + (size_t)ph * g.ntaps * g.Cin * g.Cout;
  for (int t = 0; t < g.ntaps; ++t) {
    const int iy = aqy * g.in_stride + g.dy[ph][t];
    const int ix = aqx * g.in_stride + g.dx[ph][t];
    const bool pv = a_ok && iy >= 0 && iy < g.Hin && ix >= 0 && ix < g.Win;
    const T* xrow = x + (((size_t)an * g.Hin + (pv ? iy : 0)) * g.Win + (pv ? ix : 0)) * g.Cin;
    const T* Et = Eph + (size_t)t * g.Cin * g.Cout;
    for (int c0 = 0; c0 < g.Cin; c0 += BK) {
      float av[4], bv[4];
      load_ch4<T, VEC>(xrow, c0 + a_c, g.Cin, pv, av);
      {
        const int ci = c0 + b_k;
        const int co = co0 + b_c;
        const bool bvld = ci < g.Cin && co < g.Cout;
        load_ch4<T, VEC>(Et + (size_t)(bvld ? ci : 0) * g.Cout, co, g.Cout, bvld, bv);
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) As[a_c + i][a_m] = av[i];
      *reinterpret_cast<float4*>(&Bs[b_k][b_c]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }

  const int co = co0 + tx * 4;
  if (co >= g.Cout) return;
  float bb[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (co + j < g.Cout) bb[j] = bias[co + j] * bias_mul;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= Q) continue;
    int qx = (int)(m % g.Wq);
    long long r = m / g.Wq;
    int qy = (int)(r % g.Hq);
    int n = (int)(r / g.Hq);
    int oy = qy * g.out_stride + py, ox = qx * g.out_stride + px;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = acc[i][j] + bb[j];
      if (act) v[j] = sg_lrelu(v[j]);
    }
    T* dst = y + (((size_t)n * g.Hout + oy) * g.Wout + ox) * g.Cout + co;
    if (VEC) {
      store4<T>(dst, v);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (co + j < g.Cout) dst[j] = sg_from_f<T>(v[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// wgrad: dE[ph][t][ci][co] = sum_q in[...q,t...][ci] * dy[...q,ph...][co]    (split over q into `nsplit` partials)
// ------------------------------------------------------------------------------------------------------------
template <class T, bool VEC>
__global__ void __launch_bounds__(NT) wgrad_generic_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                           float* __restrict__ partial, int nsplit, SgConvGeom g) {
  __shared__ float As[BK][BM];
  __shared__ float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int tiles_ci = (g.Cin + BM - 1) / BM;
  const int ci0 = (blockIdx.x % tiles_ci) * BM;
  const int co0 = (blockIdx.x / tiles_ci) * BN;
  const int pt = blockIdx.y;  // phase*ntaps + tap
  const int ph = pt / g.ntaps, t = pt % g.ntaps;
  const int py = ph >> 1, px = ph & 1;
  const int split = blockIdx.z;
  const long long Q = (long long)g.N * g.Hq * g.Wq;
  const long long per = ((Q + nsplit - 1) / nsplit + BK - 1) / BK * BK;
  const long long q_begin = (long long)split * per;
  const long long q_end = (q_begin + per < Q) ? (q_begin + per) : Q;

  const int l_k = tid >> 4, l_c = (tid & 15) * 4;  // load role: pixel l_k of the chunk, channel chunk l_c
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long q0 = q_begin; q0 < q_end; q0 += BK) {
    long long q = q0 + l_k;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (q < q_end) {
      int qx = (int)(q % g.Wq);
      long long r = q / g.Wq;
      int qy = (int)(r % g.Hq);
      int n = (int)(r / g.Hq);
      int iy = qy * g.in_stride + g.dy[ph][t], ix = qx * g.in_stride + g.dx[ph][t];
      bool pv = iy >= 0 && iy < g.Hin && ix >= 0 && ix < g.Win;
      if (pv) {
        const T* xr = x + (((size_t)n * g.Hin + iy) * g.Win + ix) * g.Cin;
        load_ch4<T, VEC>(xr, ci0 + l_c, g.Cin, true, av);
        int oy = qy * g.out_stride + py, ox = qx * g.out_stride + px;
        const T* dr = dy + (((size_t)n * g.Hout + oy) * g.Wout + ox) * g.Cout;
        load_ch4<T, VEC>(dr, co0 + l_c, g.Cout, true, bv);
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[l_k][l_c]) = make_float4(av[0], av[1], av[2], av[3]);
    *reinterpret_cast<float4*>(&Bs[l_k][l_c]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  float* P = partial + ((size_t)split * g.nphase * g.ntaps + pt) * g.Cin * g.Cout;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int ci = ci0 + ty * 4 + i;
    if (ci >= g.Cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = co0 + tx * 4 + j;
      if (co < g.Cout) P[(size_t)ci * g.Cout + co] = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// weight preparation: E[ph][t][a][b] = scale * sum_j coef[ph][t][j] * W[co][ci][j],  (a,b)=(ci,co) or (co,ci)
// ------------------------------------------------------------------------------------------------------------
// Tiled form: a block owns a 32 x 32 tile of (co, ci).  The parameter taps of the tile are read as 32 contiguous runs of
// 32*KK floats (coalesced whatever the output orientation is) into shared memory; every effective tap is then written
// as 32 rows of 32 consecutive elements.  (The first version -- one thread per (a,b) reading its KK taps straight from
// global memory -- ran at a fifth of the copy bandwidth: 16 us for a 512x512x3x3 weight.)
constexpr int WP_T = 32;
template <class T>
__device__ __forceinline__ void wprep_fwd_tile(const float* __restrict__ W, T* __restrict__ E, int Cout, int Cin, int KK,
                                               int npt, const float* __restrict__ coef, float scale, int transpose,
                                               int tile, float* sm_w) {
  float* s_coef = sm_w;                             // [npt*KK coef][WP_T][WP_T*KK + 1]
  float* s_w = sm_w + npt * KK;
  const int row = WP_T * KK + 1;
  for (int i = threadIdx.x; i < npt * KK; i += blockDim.x) s_coef[i] = coef[i];
  const int tiles_ci = (Cin + WP_T - 1) / WP_T;
  const int co0 = (tile / tiles_ci) * WP_T, ci0 = (tile % tiles_ci) * WP_T;
  const int nci = min(WP_T, Cin - ci0), nco = min(WP_T, Cout - co0);
  {
    // 8 independent loads in flight per thread (one load per loop trip left every trip waiting for its own DRAM round
    // trip: 20-40 us for ANY weight size)
    const int run = nci * KK, total = nco * run;
    constexpr int U = 8;
    for (int i0 = threadIdx.x; i0 < total; i0 += U * blockDim.x) {
      float v[U];
      int dsti[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * blockDim.x;
        const int c = i / run, r = i - c * run;
        dsti[u] = c * row + r;
        v[u] = i < total ? W[((size_t)(co0 + c) * Cin + ci0) * KK + r] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i0 + u * (int)blockDim.x < total) s_w[dsti[u]] = v[u] * scale;
    }
  }
  __syncthreads();
  const int A = transpose ? Cout : Cin, B = transpose ? Cin : Cout;
  const size_t AB = (size_t)A * B;
  // output element (a, b): b is the fast index in memory -> lanes walk b
  const int nb = transpose ? nci : nco, na = transpose ? nco : nci;
  const int a0 = transpose ? co0 : ci0, b0 = transpose ? ci0 : co0;
  for (int i = threadIdx.x; i < na * nb; i += blockDim.x) {
    const int bl = i % nb, al = i / nb;
    const int c = transpose ? al : bl, ci = transpose ? bl : al;
    const float* w = s_w + c * row + ci * KK;
    float wv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) wv[j] = j < KK ? w[j] : 0.f;
    T* dst = E + (size_t)(a0 + al) * B + (b0 + bl);
#pragma unroll 4
    for (int pt = 0; pt < npt; ++pt) {
      const float* cf = s_coef + pt * KK;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < KK) acc = fmaf(cf[j], wv[j], acc);
      dst[(size_t)pt * AB] = sg_from_f<T>(acc);
    }
  }
}
template <class T>
__global__ void __launch_bounds__(256) wprep_fwd_kernel(const float* __restrict__ W, T* __restrict__ E, int Cout, int Cin,
                                                        int KK, int npt, const float* __restrict__ coef, float scale,
                                                        int transpose) {
  extern __shared__ float sm_w[];
  wprep_fwd_tile<T>(W, E, Cout, Cin, KK, npt, coef, scale, transpose, blockIdx.x, sm_w);
}
// Many weights in ONE launch (the refresh of every derived bf16 copy after an optimizer step: 58 launches of 4-256 blocks
// each, ~11 us apiece, were 0.7 ms of a 30 ms step): block -> table row by binary search on the first tile of each row.
__global__ void __launch_bounds__(256) wprep_fwd_multi_kernel(const SgWprepEntry* __restrict__ tab, int n) {
  extern __shared__ float sm_w[];
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].tile0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const SgWprepEntry e = tab[lo];
  wprep_fwd_tile<__nv_bfloat16>(reinterpret_cast<const float*>(e.W), reinterpret_cast<__nv_bfloat16*>(e.E), e.Cout, e.Cin,
                                e.KK, e.npt, reinterpret_cast<const float*>(e.coef), e.scale, e.transpose,
                                (int)blockIdx.x - e.tile0, sm_w);
}

// dW[co][ci][j] (+)= scale * sum_{pt} coef[pt][j] * sum_s partial[s][pt][a][b]
// A block owns 128 consecutive (a,b) positions: phase 1, thread (pt-group, 4 positions) sums its float4 over the pixel
// splits with all loads of a split batch in flight; phase 2 combines the npt tap sums with the transposed coefficient
// table and writes dW (for transpose = 1 the block's outputs are one contiguous run of 128*KK floats).
constexpr int WB_AB = 128;   // (a,b) weight positions per CTA
constexpr int WB_PT = 16;    // tap slots per CTA (npt <= 16)
__global__ void __launch_bounds__(512)
wprep_bwd_kernel(const float* __restrict__ partial, int nsplit, float* __restrict__ dW, int Cout, int Cin, int KK,
                 int npt, const float* __restrict__ coef, float scale, int transpose, int accumulate) {
  __shared__ float s_sum[WB_PT][WB_AB + 4];
  __shared__ float s_coef[WB_PT * 16];
  const int A = transpose ? Cout : Cin, B = transpose ? Cin : Cout;
  const long long total = (long long)A * B;
  const size_t plane = (size_t)A * B;
  const size_t split_stride = (size_t)npt * plane;
  for (int i = threadIdx.x; i < npt * KK; i += blockDim.x) s_coef[i] = coef[i];
  const long long i0 = (long long)blockIdx.x * WB_AB;
  const bool vec = (plane % 4 == 0);
  // phase 1: 32 threads x float4 cover the 128 positions of one tap; 16 tap slots
  const int q = threadIdx.x & 31, pt = threadIdx.x >> 5;
  if (pt < npt) {
    const long long pos = i0 + 4 * q;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && pos + 3 < total) {
      const float* p = partial + (size_t)pt * plane + (size_t)pos;
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int sp = 0; sp < nsplit; sp += 4) {           // four independent (predicated) loads in flight
        const float4 v0 = *reinterpret_cast<const float4*>(p + (size_t)sp * split_stride);
        const float4 v1 = sp + 1 < nsplit ? *reinterpret_cast<const float4*>(p + (size_t)(sp + 1) * split_stride) : zero4;
        const float4 v2 = sp + 2 < nsplit ? *reinterpret_cast<const float4*>(p + (size_t)(sp + 2) * split_stride) : zero4;
        const float4 v3 = sp + 3 < nsplit ? *reinterpret_cast<const float4*>(p + (size_t)(sp + 3) * split_stride) : zero4;
        acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
        acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
      }
    } else {
      float* av = &acc.x;
      for (int e = 0; e < 4; ++e)
        if (pos + e < total)
          for (int sp = 0; sp < nsplit; ++sp) av[e] += partial[(size_t)sp * split_stride + (size_t)pt * plane + (size_t)(pos + e)];
    }
    s_sum[pt][4 * q] = acc.x; s_sum[pt][4 * q + 1] = acc.y; s_sum[pt][4 * q + 2] = acc.z; s_sum[pt][4 * q + 3] = acc.w;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < WB_AB * KK; t += blockDim.x) {
    const int ab2 = t / KK, j = t % KK;
    const long long i = i0 + ab2;
    if (i >= total) continue;
    float o = 0.f;
    for (int qq = 0; qq < npt; ++qq) o = fmaf(s_coef[qq * KK + j], s_sum[qq][ab2], o);
    const int b = (int)(i % B), a = (int)(i / B);
    const int co = transpose ? a : b, ci = transpose ? b : a;
    float* w = dW + ((size_t)co * Cin + ci) * KK + j;
    *w = (accumulate ? *w : 0.f) + o * scale;
  }
}

// Narrow outputs (Cout <= 4, e.g. the gradient wrt the single minibatch-stddev channel): one warp per output pixel
// and phase, lanes stride over (tap, ci), shuffle reduction.  The 64x64-tiled kernel above would run 8 CTAs with a
// 4600-step reduction each (260 us for 2.4 MFLOP).
template <class T>
__global__ void __launch_bounds__(256) conv_narrow_kernel(const T* __restrict__ x, const T* __restrict__ E,
                                                          T* __restrict__ y, const float* __restrict__ bias,
                                                          float bias_mul, int act, SgConvGeom g) {
  const int lane = threadIdx.x & 31;
  const long long Q = (long long)g.N * g.Hq * g.Wq;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= Q * g.nphase) return;              // warp-uniform
  const int ph = (int)(wid % g.nphase);
  const long long m = wid / g.nphase;
  const int qx = (int)(m % g.Wq);
  const long long r = m / g.Wq;
  const int qy = (int)(r % g.Hq);
  const int n = (int)(r / g.Hq);
  const int py = ph >> 1, px = ph & 1;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const T* Eph = E + (size_t)ph * g.ntaps * g.Cin * g.Cout;
  for (int t = 0; t < g.ntaps; ++t) {
    const int iy = qy * g.in_stride + g.dy[ph][t];
    const int ix = qx * g.in_stride + g.dx[ph][t];
    if (iy < 0 || iy >= g.Hin || ix < 0 || ix >= g.Win) continue;   // warp-uniform
    const T* xrow = x + (((size_t)n * g.Hin + iy) * g.Win + ix) * g.Cin;
    const T* Et = Eph + (size_t)t * g.Cin * g.Cout;
    for (int ci = lane; ci < g.Cin; ci += 32) {
      const float xv = sg_to_f(xrow[ci]);
      for (int co = 0; co < g.Cout; ++co) acc[co] = fmaf(xv, sg_to_f(Et[(size_t)ci * g.Cout + co]), acc[co]);
    }
  }
#pragma unroll
  for (int co = 0; co < 4; ++co) acc[co] = sg_warp_sum(acc[co]);
  if (lane < g.Cout) {
    float v = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
    if (bias != nullptr) v += bias[lane] * bias_mul;
    if (act) v = sg_lrelu(v);
    const int oy = qy * g.out_stride + py, ox = qx * g.out_stride + px;
    y[(((size_t)n * g.Hout + oy) * g.Wout + ox) * g.Cout + lane] = sg_from_f<T>(v);
  }
}

bool geom_ok(const SgConvGeom* g) {
  return g && g->N > 0 && g->Hq > 0 && g->Wq > 0 && g->Cin > 0 && g->Cout > 0 && g->ntaps >= 1 && g->ntaps <= 16 &&
         (g->nphase == 1 || g->nphase == 4) && g->in_stride >= 1 && g->out_stride >= 1;
}

}  // namespace

SG_EXPORT int sg_conv_generic(const void* x, const void* E, void* y, const float* bias, float bias_mul, int act,
                              int dtype, const SgConvGeom* g, void* stream) {
  SG_REQUIRE(geom_ok(g), "sg_conv_generic: bad geometry");
  long long Q = (long long)g->N * g->Hq * g->Wq;
  dim3 grid(sg_cdiv(Q, BM), sg_cdiv(g->Cout, BN), g->nphase);
  bool vec = (g->Cin % 4 == 0) && (g->Cout % 4 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  if (g->Cout <= 4 && Q * g->nphase <= (1ll << 22)) {      // narrow output, small map: warp-per-pixel reduction
    const long long warps = Q * g->nphase;
    SG_DISPATCH(dtype, {
      conv_narrow_kernel<T><<<sg_cdiv(warps * 32, 256), 256, 0, st>>>((const T*)x, (const T*)E, (T*)y, bias, bias_mul, act, *g);
    });
    return sg_check_launch("sg_conv_generic/narrow");
  }
  SG_DISPATCH(dtype, {
    if (vec) conv_generic_kernel<T, true><<<grid, NT, 0, st>>>((const T*)x, (const T*)E, (T*)y, bias, bias_mul, act, *g);
    else conv_generic_kernel<T, false><<<grid, NT, 0, st>>>((const T*)x, (const T*)E, (T*)y, bias, bias_mul, act, *g);
  });
  return sg_check_launch("sg_conv_generic");
}

SG_EXPORT int sg_wgrad_generic(const void* x, const void* dy, float* partial, int nsplit, int dtype,
                               const SgConvGeom* g, void* stream) {
  SG_REQUIRE(geom_ok(g) && nsplit >= 1, "sg_wgrad_generic: bad geometry");
  dim3 grid(sg_cdiv(g->Cin, BM) * sg_cdiv(g->Cout, BN), g->nphase * g->ntaps, nsplit);
  bool vec = (g->Cin % 4 == 0) && (g->Cout % 4 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  SG_DISPATCH(dtype, {
    if (vec) wgrad_generic_kernel<T, true><<<grid, NT, 0, st>>>((const T*)x, (const T*)dy, partial, nsplit, *g);
    else wgrad_generic_kernel<T, false><<<grid, NT, 0, st>>>((const T*)x, (const T*)dy, partial, nsplit, *g);
  });
  return sg_check_launch("sg_wgrad_generic");
}

SG_EXPORT int sg_wgrad_splits(const SgConvGeom* g) {
  // enough CTAs to fill 148 SMs a few times over, at least 256 pixels per split
  if (!geom_ok(g)) return 1;
  long long Q = (long long)g->N * g->Hq * g->Wq;
  long long tiles = (long long)sg_cdiv(g->Cin, BM) * sg_cdiv(g->Cout, BN) * g->nphase * g->ntaps;
  long long want = (148LL * 4 + tiles - 1) / tiles;
  long long maxs = Q / 256;
  if (maxs < 1) maxs = 1;
  if (want > maxs) want = maxs;
  if (want > 256) want = 256;
  if (want < 1) want = 1;
  return (int)want;
}

SG_EXPORT int sg_wprep_fwd(const float* W, void* E, int dtype, int Cout, int Cin, int KK, int npt, const float* coef,
                           float scale, int transpose, void* stream) {
  SG_REQUIRE(KK >= 1 && KK <= 16 && npt >= 1, "sg_wprep_fwd: bad sizes");
  SG_REQUIRE(npt <= 64, "sg_wprep_fwd: too many effective taps");
  cudaStream_t st = (cudaStream_t)stream;
  SG_DISPATCH(dtype, {
    const int tiles = sg_cdiv(Cout, WP_T) * sg_cdiv(Cin, WP_T);
    const size_t smem = ((size_t)npt * KK + (size_t)WP_T * (WP_T * KK + 1)) * sizeof(float);
    static bool attr_done = false;   // KK = 16 needs 66 KB
    if (!attr_done) {
      cudaFuncSetAttribute(wprep_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      attr_done = true;
    }
    wprep_fwd_kernel<T><<<tiles, 256, smem, st>>>(W, (T*)E, Cout, Cin, KK, npt, coef, scale, transpose);
  });
  return sg_check_launch("sg_wprep_fwd");
}

// table: n rows of SgWprepEntry in device memory (bf16 outputs; tile0 = running sum of ceil(Cout/32)*ceil(Cin/32));
// total_tiles = the sum over all rows; max_smem_floats = max over rows of npt*KK + 32*(32*KK + 1).
SG_EXPORT int sg_wprep_fwd_multi(const void* table, int n, int total_tiles, int max_smem_floats, void* stream) {
  if (n <= 0 || total_tiles <= 0) return 0;
  SG_REQUIRE(table != nullptr && max_smem_floats > 0 && (size_t)max_smem_floats * sizeof(float) <= 96 * 1024,
             "sg_wprep_fwd_multi: bad arguments");
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(wprep_fwd_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  });
  wprep_fwd_multi_kernel<<<total_tiles, 256, (size_t)max_smem_floats * sizeof(float), (cudaStream_t)stream>>>(
      reinterpret_cast<const SgWprepEntry*>(table), n);
  return sg_check_launch("sg_wprep_fwd_multi");
}

SG_EXPORT int sg_wprep_bwd(const float* partial, int nsplit, float* dW, int Cout, int Cin, int KK, int npt,
                           const float* coef, float scale, int transpose, int accumulate, void* stream) {
  SG_REQUIRE(KK >= 1 && KK <= 16 && npt >= 1 && npt <= WB_PT && nsplit >= 1, "sg_wprep_bwd: bad sizes");
  long long total = (long long)Cout * Cin;
  wprep_bwd_kernel<<<sg_cdiv(total, WB_AB), 32 * WB_PT, 0, (cudaStream_t)stream>>>(partial, nsplit, dW, Cout, Cin, KK, npt,
                                                                             coef, scale, transpose, accumulate);
  return sg_check_launch("sg_wprep_bwd");
}
