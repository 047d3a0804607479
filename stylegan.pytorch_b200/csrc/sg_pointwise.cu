// Bandwidth-bound NHWC kernels: blur, axpby/lerp, LeakyReLU backward, casts, per-channel reductions, from/to-RGB
// pointwise convolutions (NCHW fp32 image side <-> NHWC feature side), image pooling, dlatent assembly.
#include "sg_common.cuh"
#include "sg_fast.cuh"
#include "../../include/sgb200.h"

namespace {

__device__ __forceinline__ float pick4(const float (&a)[4], int k) {
  return k == 0 ? a[0] : (k == 1 ? a[1] : (k == 2 ? a[2] : a[3]));
}

// ---------------------------------------------------------------------------------------------------------
// depthwise [1,2,1]x[1,2,1]/16 blur, zero pad 1, optional +bias[c]*bias_mul.   reference CustomLayers.py:251-276
// self-adjoint, so the same kernel is its own backward (without bias).
// ---------------------------------------------------------------------------------------------------------
// Each thread owns one 16-byte channel vector of one image column and walks BLUR_STRIP output rows, keeping a 3-row
// window of horizontal [1,2,1] sums in registers (3.75 loads per output instead of 9).  The STRIP+2 input rows are
// fetched in batches of BLUR_BATCH rows x 3 columns = 15 independent 16-byte loads before any arithmetic: with one
// row (3 loads) in flight per thread the kernel ran at 2.5 TB/s, latency-bound.
constexpr int BLUR_STRIP = 8;
constexpr int BLUR_BATCH = 5;
static_assert((BLUR_STRIP + 2) % BLUR_BATCH == 0, "strip + halo must be a whole number of batches");
template <class T>
__global__ void __launch_bounds__(128) blur3_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                    const float* __restrict__ bias, float bias_mul, int N, int H,
                                                    int W, int C) {
  using V = SgVec<T>;
  const int CV = C / V::N;
  const int strips = (H + BLUR_STRIP - 1) / BLUR_STRIP;
  const long long total = (long long)N * strips * W * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    long long p = i / CV;
    int w = (int)(p % W);
    long long r = p / W;
    int st = (int)(r % strips);
    int n = (int)(r / strips);
    const int h0 = st * BLUR_STRIP;
    float bb[V::N];
#pragma unroll
    for (int k = 0; k < V::N; ++k) bb[k] = bias ? bias[cv * V::N + k] * bias_mul : 0.f;
    const bool has_l = w > 0, has_r = w + 1 < W;
    const T* col = x + ((size_t)n * H * W + w) * C + cv * V::N;    // (n, row 0, w, cv)
    const int off_l = has_l ? -C : 0, off_r = has_r ? C : 0;      // clamped: every load below is unconditional,
    float r0[V::N] = {}, r1[V::N] = {};                           // so ptxas can issue a whole batch back to back
#pragma unroll
    for (int b = 0; b < (BLUR_STRIP + 2) / BLUR_BATCH; ++b) {
      V raw[BLUR_BATCH][3];
#pragma unroll
      for (int j = 0; j < BLUR_BATCH; ++j) {
        int hh = h0 - 1 + b * BLUR_BATCH + j;
        hh = hh < 0 ? 0 : (hh >= H ? H - 1 : hh);
        const T* row = col + (size_t)hh * W * C;
        raw[j][0].load(row + off_l);
        raw[j][1].load(row);
        raw[j][2].load(row + off_r);
      }
#pragma unroll
      for (int j = 0; j < BLUR_BATCH; ++j) {
        const int ri = b * BLUR_BATCH + j;      // input row h0 - 1 + ri
        const int hh = h0 - 1 + ri;
        const bool ok = hh >= 0 && hh < H;
        float hs[V::N];
#pragma unroll
        for (int k = 0; k < V::N; ++k) {
          const float l = has_l ? raw[j][0].get(k) : 0.f, r = has_r ? raw[j][2].get(k) : 0.f;
          const float t = (2.f * raw[j][1].get(k) + l) + r;
          hs[k] = ok ? t : 0.f;
        }
        if (ri >= 2) {
          const int h = h0 + ri - 2;
          if (h < H) {
            V o;
#pragma unroll
            for (int k = 0; k < V::N; ++k) o.set(k, (r0[k] + 2.f * r1[k] + hs[k]) * (1.f / 16.f) + bb[k]);
            o.store(y + (((size_t)n * H + h) * W + w) * C + cv * V::N);
          }
        }
#pragma unroll
        for (int k = 0; k < V::N; ++k) { r0[k] = r1[k]; r1[k] = hs[k]; }
      }
    }
  }
}

// scalar fallback for C not a multiple of the vector width
template <class T>
__global__ void blur3_scalar_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ bias,
                                    float bias_mul, int N, int H, int W, int C) {
  const long long total = (long long)N * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long p = i / C;
    int w = (int)(p % W);
    long long r = p / W;
    int h = (int)(r % H);
    int n = (int)(r / H);
    float acc = 0.f;
    for (int dy = -1; dy <= 1; ++dy) {
      int hh = h + dy;
      if (hh < 0 || hh >= H) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        int ww = w + dx;
        if (ww < 0 || ww >= W) continue;
        float kw = (dy == 0 ? 2.f : 1.f) * (dx == 0 ? 2.f : 1.f) * (1.f / 16.f);
        acc = fmaf(kw, sg_to_f(x[(((size_t)n * H + hh) * W + ww) * C + c]), acc);
      }
    }
    y[i] = sg_from_f<T>(acc + (bias ? bias[c] * bias_mul : 0.f));
  }
}

// ---------------------------------------------------------------------------------------------------------
// out = a*x + b*y   (y may be null).  Used for fade-in lerps (GAN.py:202,427,586), scaling, EMA-style updates.
// ---------------------------------------------------------------------------------------------------------
template <class T>
__global__ void axpby_kernel(const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ out, float a, float b,
                             long long n, const float* __restrict__ ab_dev = nullptr) {
  using V = SgVec<T>;
  if (ab_dev) { a = ab_dev[0]; b = y ? ab_dev[1] : 0.f; }      // coefficients from device memory (CUDA-graph replay)
  const long long nv = n / V::N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    V vx, vy, vo;
    vx.load(x + i * V::N);
    if (y) vy.load(y + i * V::N); else vy.zero();
#pragma unroll
    for (int k = 0; k < V::N; ++k) vo.set(k, a * vx.get(k) + b * vy.get(k));
    vo.store(out + i * V::N);
  }
  // tail
  for (long long i = nv * V::N + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = sg_from_f<T>(a * sg_to_f(x[i]) + (y ? b * sg_to_f(y[i]) : 0.f));
}

// gx = g * (ref > 0 ? 1 : slope): backward of LeakyReLU given its OUTPUT (or input; same sign).
template <class T>
__global__ void lrelu_bwd_kernel(const T* __restrict__ g, const T* __restrict__ ref, T* __restrict__ gx,
                                 long long n) {
  using V = SgVec<T>;
  const long long nv = n / V::N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    V vg, vr, vo;
    vg.load(g + i * V::N);
    vr.load(ref + i * V::N);
#pragma unroll
    for (int k = 0; k < V::N; ++k) vo.set(k, vr.get(k) > 0.f ? vg.get(k) : SG_LRELU_SLOPE * vg.get(k));
    vo.store(gx + i * V::N);
  }
  for (long long i = nv * V::N + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    gx[i] = sg_from_f<T>(sg_to_f(ref[i]) > 0.f ? sg_to_f(g[i]) : SG_LRELU_SLOPE * sg_to_f(g[i]));
}

// gx = g * lrelu'(ref) AND part[blk][c] = sum over the block's pixels of the (rounded) gx: the backward of a fused
// conv + bias + LeakyReLU needs both, and the bias gradient as a second pass re-read all of gx (channel_sum_ones_kernel).
// Work split and partial layout of channel_sum_ones_kernel (a thread owns one 16-byte channel vector, strides over pixels).
template <class T>
__global__ void __launch_bounds__(256) lrelu_bwd_bias_kernel(const T* __restrict__ g, const T* __restrict__ ref,
                                                             T* __restrict__ gx, long long P, int C,
                                                             float* __restrict__ part, int pix_per_blk) {
  using V = SgVec<T>;
  extern __shared__ float sm[];                 // [PR][C]
  const int CV = C / V::N;
  const int PR = blockDim.x / CV;
  const int cv = threadIdx.x % CV, pr = threadIdx.x / CV;
  const long long p0 = (long long)blockIdx.x * pix_per_blk;
  long long p1 = p0 + pix_per_blk;
  if (p1 > P) p1 = P;
  float acc[V::N];
#pragma unroll
  for (int k = 0; k < V::N; ++k) acc[k] = 0.f;
  if (pr < PR) {
    constexpr int U = 4;                          // pixels in flight per thread (two loads each)
    const size_t step = (size_t)PR * C;
    long long pb = p0 + pr;
    size_t off = (size_t)pb * C + cv * V::N;
    for (; pb + (long long)(U - 1) * PR < p1; pb += (long long)U * PR, off += U * step) {
      V vg[U], vr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { vg[u].load(g + off + u * step); vr[u].load(ref + off + u * step); }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        V vo;
#pragma unroll
        for (int k = 0; k < V::N; ++k) vo.set(k, vr[u].get(k) > 0.f ? vg[u].get(k) : SG_LRELU_SLOPE * vg[u].get(k));
        vo.store(gx + off + u * step);
#pragma unroll
        for (int k = 0; k < V::N; ++k) acc[k] += vo.get(k);
      }
    }
    for (; pb < p1; pb += PR, off += step) {
      V vg, vr, vo;
      vg.load(g + off); vr.load(ref + off);
#pragma unroll
      for (int k = 0; k < V::N; ++k) vo.set(k, vr.get(k) > 0.f ? vg.get(k) : SG_LRELU_SLOPE * vg.get(k));
      vo.store(gx + off);
#pragma unroll
      for (int k = 0; k < V::N; ++k) acc[k] += vo.get(k);
    }
#pragma unroll
    for (int k = 0; k < V::N; ++k) sm[(size_t)pr * C + cv * V::N + k] = acc[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int r = 0; r < PR; ++r) t += sm[(size_t)r * C + c];
    part[(size_t)blockIdx.x * C + c] = t;
  }
}

template <class TI, class TO>
__global__ void cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = sg_from_f<TO>(sg_to_f(x[i]));
}

// ---------------------------------------------------------------------------------------------------------
// Weighted per-channel reduction over pixels:  part[blk][j][c] = sum_{p in blk} wt_j(p) * x[p][c]
//   rows j = 0..J-1 use wt[(n*J + j)*HW + hw]  (p = n*HW + hw);  if `ones` an extra last row uses weight 1.
// Stage 2 sums the partials deterministically.  Used for bias / noise-weight / from_rgb / to_rgb gradients.
// ---------------------------------------------------------------------------------------------------------
constexpr int CS_ROWS = 4;  // max J + ones
template <class T>
__global__ void channel_sum_kernel(const T* __restrict__ x, const float* __restrict__ wt, int J, int ones,
                                   long long P, int HW, int C, float* __restrict__ part, int pix_per_blk) {
  extern __shared__ float sm[];  // [rows][R][C] reduced over R at the end
  const int R = blockDim.y;
  const int rows = J + (ones ? 1 : 0);
  const long long p0 = (long long)blockIdx.x * pix_per_blk;
  long long p1 = p0 + pix_per_blk;
  if (p1 > P) p1 = P;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc[CS_ROWS];
#pragma unroll
    for (int j = 0; j < CS_ROWS; ++j) acc[j] = 0.f;
    for (long long p = p0 + threadIdx.y; p < p1; p += R) {
      float v = sg_to_f(x[(size_t)p * C + c]);
      long long n = p / HW;
      int hw = (int)(p - n * HW);
#pragma unroll
      for (int j = 0; j < CS_ROWS; ++j) {
        if (j < J) acc[j] = fmaf(wt[((size_t)n * J + j) * HW + hw], v, acc[j]);
      }
#pragma unroll
      for (int j = 0; j < CS_ROWS; ++j)
        if (ones && j == J) acc[j] += v;
    }
    for (int j = 0; j < rows; ++j) sm[((size_t)j * R + threadIdx.y) * C + c] = pick4(acc, j);
  }
  __syncthreads();
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
  for (int i = tid; i < rows * C; i += nthr) {
    int j = i / C, c = i % C;
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += sm[((size_t)j * R + r) * C + c];
    part[((size_t)blockIdx.x * rows + j) * C + c] = s;
  }
}

// Vectorised variant (C a multiple of the 16-byte vector width): a thread owns one 16-byte channel vector of a pixel
// row and strides over pixels; per-thread partials for all rows live in registers; one smem reduction per block.
template <class T>
__global__ void __launch_bounds__(256) channel_sum_vec_kernel(const T* __restrict__ x, const float* __restrict__ wt,
                                                              int J, int ones, long long P, int HW, int C,
                                                              float* __restrict__ part, int pix_per_blk) {
  using V = SgVec<T>;
  extern __shared__ float sm[];                 // [rows][PR][C]
  const int CV = C / V::N;
  const int PR = blockDim.x / CV;               // pixel rows per iteration
  const int cv = threadIdx.x % CV, pr = threadIdx.x / CV;
  const int rows = J + (ones ? 1 : 0);
  const long long p0 = (long long)blockIdx.x * pix_per_blk;
  long long p1 = p0 + pix_per_blk;
  if (p1 > P) p1 = P;
  float acc[CS_ROWS][V::N];
#pragma unroll
  for (int j = 0; j < CS_ROWS; ++j)
#pragma unroll
    for (int k = 0; k < V::N; ++k) acc[j][k] = 0.f;
  if (pr < PR) {
    constexpr int U = 4;                          // pixels in flight per thread
    for (long long pb = p0 + pr; pb < p1; pb += (long long)U * PR) {
      V v[U];
      float w[U][CS_ROWS];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = pb + (long long)u * PR;
        if (p < p1) {
          v[u].load(x + (size_t)p * C + cv * V::N);
          const unsigned n = (unsigned)p / (unsigned)HW;          // P < 2^31 (launcher): 32-bit division
          const int hw = (int)((unsigned)p - n * (unsigned)HW);
#pragma unroll
          for (int j = 0; j < CS_ROWS; ++j)
            w[u][j] = (j < J) ? wt[((size_t)n * J + j) * HW + hw] : ((ones && j == J) ? 1.f : 0.f);
        } else {
          v[u].zero();
#pragma unroll
          for (int j = 0; j < CS_ROWS; ++j) w[u][j] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < V::N; ++k) {
          float f = v[u].get(k);
#pragma unroll
          for (int j = 0; j < CS_ROWS; ++j) acc[j][k] = fmaf(w[u][j], f, acc[j][k]);
        }
    }
#pragma unroll
    for (int j = 0; j < CS_ROWS; ++j)
      if (j < rows) {
#pragma unroll
        for (int k = 0; k < V::N; ++k) sm[((size_t)j * PR + pr) * C + cv * V::N + k] = acc[j][k];
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows * C; i += blockDim.x) {
    int j = i / C, c = i % C;
    float s = 0.f;
    for (int r = 0; r < PR; ++r) s += sm[((size_t)j * PR + r) * C + c];
    part[((size_t)blockIdx.x * rows + j) * C + c] = s;
  }
}

// Plain per-channel sum (no pixel weights): the bias gradient of every convolution.  Same work split and partial
// layout as channel_sum_vec_kernel with rows = 1, but 8 independent 16-byte loads per thread and one add per element
// (the weighted kernel spends 4 FMAs + index arithmetic per element and ran at 1.9 TB/s).
template <class T>
__global__ void __launch_bounds__(256) channel_sum_ones_kernel(const T* __restrict__ x, long long P, int C,
                                                               float* __restrict__ part, int pix_per_blk) {
  using V = SgVec<T>;
  extern __shared__ float sm[];                 // [PR][C]
  const int CV = C / V::N;
  const int PR = blockDim.x / CV;
  const int cv = threadIdx.x % CV, pr = threadIdx.x / CV;
  const long long p0 = (long long)blockIdx.x * pix_per_blk;
  long long p1 = p0 + pix_per_blk;
  if (p1 > P) p1 = P;
  float acc[V::N];
#pragma unroll
  for (int k = 0; k < V::N; ++k) acc[k] = 0.f;
  if (pr < PR) {
    constexpr int U = 8;                          // the tree below is written for 8
    const size_t step = (size_t)PR * C;           // elements between this thread's consecutive pixels
    long long pb = p0 + pr;
    const T* ptr = x + (size_t)pb * C + cv * V::N;
    for (; pb + (long long)(U - 1) * PR < p1; pb += (long long)U * PR, ptr += U * step) {
      V v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u].load(ptr + u * step);      // U independent loads, no predicates
      // pairwise tree: a serial acc += v[u] chain lets ptxas sink load u to just before its add (3 loads in flight)
#pragma unroll
      for (int k = 0; k < V::N; ++k) {
        const float a = v[0].get(k) + v[1].get(k), b = v[2].get(k) + v[3].get(k);
        const float c = v[4].get(k) + v[5].get(k), d = v[6].get(k) + v[7].get(k);
        acc[k] += (a + b) + (c + d);
      }
    }
    for (; pb < p1; pb += PR, ptr += step) {
      V v;
      v.load(ptr);
#pragma unroll
      for (int k = 0; k < V::N; ++k) acc[k] += v.get(k);
    }
#pragma unroll
    for (int k = 0; k < V::N; ++k) sm[(size_t)pr * C + cv * V::N + k] = acc[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < PR; ++r) s += sm[(size_t)r * C + c];
    part[(size_t)blockIdx.x * C + c] = s;
  }
}

// out[i] = (accumulate ? out[i] : 0) + scale * sum_b part[b][i]: one warp per output element, lanes stride over the
// partial blocks (one thread walking ~600 partials serially cost 35 us per call).
__global__ void __launch_bounds__(128) partial_reduce_kernel(const float* __restrict__ part, int nblk, int M,
                                                             float* __restrict__ out, float scale, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= M) return;                           // warp-uniform
  float s = 0.f;
  for (int b = lane; b < nblk; b += 32) s += part[(size_t)b * M + i];
  s = sg_warp_sum(s);
  if (lane == 0) out[i] = (accumulate ? out[i] : 0.f) + s * scale;
}

// ---------------------------------------------------------------------------------------------------------
// from_rgb: img fp32 NCHW [N,K,H,W] -> y NHWC T [N,H,W,C]:  y = act(scale * sum_k W[c][k] img[k] + b[c]*bmul)
// reference GAN.py:358,377 (1x1 EqualizedConv2d) + LeakyReLU.  K <= 4.
// ---------------------------------------------------------------------------------------------------------
// vectorised: one thread = one pixel x 16-byte channel vector.  The launch makes gridDim*blockDim a multiple of CV,
// so a thread keeps the same channel vector for its whole grid-stride loop and holds its weights in registers.
template <class T>
__global__ void rgb_in_vec_kernel(const float* __restrict__ img, const float* __restrict__ Wt,
                                  const float* __restrict__ bias, float scale, float bias_mul, int act,
                                  T* __restrict__ y, int N, int HW, int K, int C) {
  using V = SgVec<T>;
  const int CV = C / V::N;
  const int total = N * HW * CV;                 // < 2^31 checked by the launcher
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = gridDim.x * blockDim.x;     // a multiple of CV
  const int cv = i0 % CV;
  float wr[V::N][4], br[V::N];
#pragma unroll
  for (int e = 0; e < V::N; ++e) {
    const int c = cv * V::N + e;
    br[e] = bias ? bias[c] * bias_mul : 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) wr[e][k] = (k < K) ? Wt[c * K + k] * scale : 0.f;
  }
  // two pixels per iteration: the image loads of both are issued before either result is formed
  for (int i = i0; i < total; i += 2 * stride) {
    const bool ok2 = i + stride < total;
    const int pa = i / CV, pb = (ok2 ? i + stride : i) / CV;
    const int na = pa / HW, nb = pb / HW;
    const int ha = pa - na * HW, hb = pb - nb * HW;
    float xa[4] = {0.f, 0.f, 0.f, 0.f}, xb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < K) {
        xa[k] = img[((size_t)na * K + k) * HW + ha];
        xb[k] = img[((size_t)nb * K + k) * HW + hb];
      }
    V oa, ob;
#pragma unroll
    for (int e = 0; e < V::N; ++e) {
      float sa = br[e], sb = br[e];
#pragma unroll
      for (int k = 0; k < 4; ++k) { sa = fmaf(wr[e][k], xa[k], sa); sb = fmaf(wr[e][k], xb[k], sb); }
      if (act) { sa = sg_lrelu(sa); sb = sg_lrelu(sb); }
      oa.set(e, sa);
      ob.set(e, sb);
    }
    oa.store(y + (size_t)pa * C + cv * V::N);
    if (ok2) ob.store(y + (size_t)pb * C + cv * V::N);
  }
}

template <class T>
__global__ void rgb_in_kernel(const float* __restrict__ img, const float* __restrict__ Wt, const float* __restrict__ bias,
                              float scale, float bias_mul, int act, T* __restrict__ y, int N, int HW, int K, int C) {
  const long long total = (long long)N * HW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long p = i / C;
    long long n = p / HW;
    int hw = (int)(p - n * HW);
    float s = 0.f;
    for (int k = 0; k < K; ++k) s = fmaf(Wt[c * K + k], img[((size_t)n * K + k) * HW + hw], s);
    s = s * scale + (bias ? bias[c] * bias_mul : 0.f);
    if (act) s = sg_lrelu(s);
    y[i] = sg_from_f<T>(s);
  }
}

// adjoint wrt the image: gimg[n,k,hw] = scale * sum_c g[n,hw,c] * W[c][k]
// a group of G = min(32, C/VEC) lanes owns one pixel (16-byte loads), xor-shuffle reduction inside the group.
template <class T>
__global__ void __launch_bounds__(256) rgb_in_dgrad_kernel(const T* __restrict__ g, const float* __restrict__ Wt,
                                                           float scale, float* __restrict__ gimg, int N, int HW,
                                                           int K, int C, int G) {
  using V = SgVec<T>;
  constexpr int U = 4;                           // pixels in flight per lane group
  const int P = N * HW;                          // < 2^31 checked by the launcher
  const int lane = threadIdx.x & 31;
  const int sub = lane / G, l = lane % G, per_warp = 32 / G;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int CV = C / V::N;
  const int lc = l < CV ? l : CV - 1;
  // weights of this lane's first channel vector in registers (the only one when CV <= G, the common case)
  float wreg[V::N][4];
#pragma unroll
  for (int e = 0; e < V::N; ++e)
#pragma unroll
    for (int k = 0; k < 4; ++k) wreg[e][k] = (l < CV && k < K) ? Wt[(l * V::N + e) * K + k] : 0.f;
  for (long long q0 = (long long)warp * per_warp * U; q0 < P; q0 += (long long)nwarps * per_warp * U) {
    V v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = q0 + u * per_warp + sub;   // clamped + unconditional: lanes >= CV carry zero weights and
      v[u].load(g + (size_t)(p < P ? p : P - 1) * C + lc * V::N);   // pixels >= P are never written
    }
    float acc[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[u][k] = 0.f;
#pragma unroll
      for (int e = 0; e < V::N; ++e) {
        const float f = v[u].get(e);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[u][k] = fmaf(f, wreg[e][k], acc[u][k]);
      }
    }
    if (CV > G) {                                // wide layers (C > 32 vectors): remaining vectors, weights from L1
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = q0 + u * per_warp + sub;
        if (p < P)
          for (int cv = l + G; cv < CV; cv += G) {
            V t;
            t.load(g + (size_t)p * C + cv * V::N);
#pragma unroll
            for (int e = 0; e < V::N; ++e) {
              const float f = t.get(e);
              const float* wr = Wt + (cv * V::N + e) * K;
              for (int k = 0; k < K; ++k) acc[u][k] = fmaf(f, wr[k], acc[u][k]);
            }
          }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        for (int o = G >> 1; o > 0; o >>= 1) acc[u][k] += __shfl_xor_sync(0xffffffffu, acc[u][k], o);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = q0 + u * per_warp + sub;
      if (p < P && l < K) {
        const int n = (int)p / HW;
        const int hw = (int)p - n * HW;
        gimg[((size_t)n * K + l) * HW + hw] = pick4(acc[u], l) * scale;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// to_rgb: x NHWC T [N,h,w,C] -> out fp32 NCHW [N,K,h*up,w*up]:
//   out = beta*out + alpha*(scale * sum_c W[k][c] x[c] + b[k]*bmul), nearest-upsampled by `up`.
// reference GAN.py:157,167 (1x1 conv, gain 1), GAN.py:199-202 (temporaryUpsampler + fade-in lerp).
// ---------------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) rgb_out_kernel(const T* __restrict__ x, const float* __restrict__ Wt,
                                                      const float* __restrict__ bias, float scale, float bias_mul,
                                                      float alpha, float beta, float* __restrict__ out, int N, int h,
                                                      int w, int C, int K, int up, int G) {
  using V = SgVec<T>;
  constexpr int U = 4;                           // pixels in flight per lane group
  const int P = N * h * w;                       // < 2^31 checked by the launcher
  const int lane = threadIdx.x & 31;
  const int sub = lane / G, l = lane % G, per_warp = 32 / G;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int H = h * up, Wd = w * up, CV = C / V::N;
  const int lc = l < CV ? l : CV - 1;
  float wreg[V::N][4];                           // this lane's first channel vector of W[k][c]
#pragma unroll
  for (int e = 0; e < V::N; ++e)
#pragma unroll
    for (int k = 0; k < 4; ++k) wreg[e][k] = (l < CV && k < K) ? Wt[k * C + l * V::N + e] : 0.f;
  for (long long q0 = (long long)warp * per_warp * U; q0 < P; q0 += (long long)nwarps * per_warp * U) {
    V v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = q0 + u * per_warp + sub;   // clamped + unconditional: lanes >= CV carry zero weights and
      v[u].load(x + (size_t)(p < P ? p : P - 1) * C + lc * V::N);   // pixels >= P are never written
    }
    float acc[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[u][k] = 0.f;
#pragma unroll
      for (int e = 0; e < V::N; ++e) {
        const float f = v[u].get(e);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[u][k] = fmaf(f, wreg[e][k], acc[u][k]);
      }
    }
    if (CV > G) {                                // wide layers: remaining vectors, weights from L1
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = q0 + u * per_warp + sub;
        if (p < P)
          for (int cv = l + G; cv < CV; cv += G) {
            V t;
            t.load(x + (size_t)p * C + cv * V::N);
#pragma unroll
            for (int e = 0; e < V::N; ++e) {
              const float f = t.get(e);
              for (int k = 0; k < K; ++k) acc[u][k] = fmaf(f, Wt[k * C + cv * V::N + e], acc[u][k]);
            }
          }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        for (int o = G >> 1; o > 0; o >>= 1) acc[u][k] += __shfl_xor_sync(0xffffffffu, acc[u][k], o);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long pl = q0 + u * per_warp + sub;
      if (pl < P) {
        const int p = (int)pl;
        const int xw = p % w;
        const int r = p / w;
        const int yh = r % h;
        const int n = r / h;
        // the group's lanes spread over (k, a, b)
        for (int j = l; j < K * up * up; j += G) {
          int k = j / (up * up), ab = j % (up * up), a = ab / up, b = ab % up;
          float val = alpha * (pick4(acc[u], k) * scale + (bias ? bias[k] * bias_mul : 0.f));
          size_t o = (((size_t)n * K + k) * H + (yh * up + a)) * Wd + (xw * up + b);
          out[o] = (beta != 0.f ? beta * out[o] : 0.f) + val;
        }
      }
    }
  }
}

// adjoint wrt x: gx[n,y,x,c] = alpha*scale * sum_k W[k][c] * sum_{a,b<up} g[n,k,y*up+a,x*up+b]
template <class T>
__global__ void rgb_out_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ Wt, float coef,
                                     T* __restrict__ gx, int N, int h, int w, int C, int K, int up) {
  using V = SgVec<T>;
  const int CV = C / V::N;
  const long long total = (long long)N * h * w * CV;
  const int H = h * up, Wd = w * up;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    long long p = i / CV;
    int xw = (int)(p % w);
    long long r = p / w;
    int yh = (int)(r % h);
    int n = (int)(r / h);
    float gs[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < K; ++k)
      for (int a = 0; a < up; ++a)
        for (int b = 0; b < up; ++b) gs[k] += g[(((size_t)n * K + k) * H + (yh * up + a)) * Wd + (xw * up + b)];
    V o;
#pragma unroll
    for (int e = 0; e < V::N; ++e) {
      float s_ = 0.f;
      for (int k = 0; k < K; ++k) s_ = fmaf(Wt[k * C + cv * V::N + e], gs[k], s_);
      o.set(e, s_ * coef);
    }
    o.store(gx + (size_t)p * C + cv * V::N);
  }
}

// same, launched with gridDim*blockDim a multiple of CV: a thread keeps one channel vector for its whole grid-stride
// loop, so W*coef lives in registers and the index arithmetic is 32-bit.
template <class T>
__global__ void rgb_out_dgrad_vec_kernel(const float* __restrict__ g, const float* __restrict__ Wt, float coef,
                                         T* __restrict__ gx, int N, int h, int w, int C, int K, int up) {
  using V = SgVec<T>;
  const int CV = C / V::N;
  const int total = N * h * w * CV;              // < 2^31 checked by the launcher
  const int H = h * up, Wd = w * up;
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = gridDim.x * blockDim.x;     // a multiple of CV
  const int cv = i0 % CV;
  float wr[V::N][4];
#pragma unroll
  for (int e = 0; e < V::N; ++e)
#pragma unroll
    for (int k = 0; k < 4; ++k) wr[e][k] = (k < K) ? Wt[k * C + cv * V::N + e] * coef : 0.f;
  for (int i = i0; i < total; i += stride) {
    const int p = i / CV;
    const int xw = p % w;
    const int r = p / w;
    const int yh = r % h;
    const int n = r / h;
    float gs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < K) {
        const float* gp = g + (((size_t)n * K + k) * H + (size_t)yh * up) * Wd + (size_t)xw * up;
        float t = gp[0];
        if (up == 2) t += gp[1] + gp[Wd] + gp[Wd + 1];
        gs[k] = t;
      }
    V o;
#pragma unroll
    for (int e = 0; e < V::N; ++e) {
      float s_ = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) s_ = fmaf(wr[e][k], gs[k], s_);
      o.set(e, s_);
    }
    o.store(gx + (size_t)p * C + cv * V::N);
  }
}

// ---------------------------------------------------------------------------------------------------------
// fp32 NCHW image helpers: box pooling / nearest up-sampling (adjoint pair).   reference GAN.py:382,575-586
// ---------------------------------------------------------------------------------------------------------
__global__ void pool_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int Ho,
                                 int Wo, int f, float scale) {
  const long long total = planes * Ho * Wo;
  const int Wi = Wo * f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % Wo);
    long long r = i / Wo;
    int y = (int)(r % Ho);
    long long pl = r / Ho;
    const float* src = in + (pl * Ho * f + (long long)y * f) * Wi + (long long)x * f;
    float s = 0.f;
    for (int a = 0; a < f; ++a)
      for (int b = 0; b < f; ++b) s += src[(long long)a * Wi + b];
    out[i] = s * scale;
  }
}

__global__ void upsample_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int Hi,
                                     int Wi, int f, float scale) {
  const int Ho = Hi * f, Wo = Wi * f;
  const long long total = planes * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % Wo);
    long long r = i / Wo;
    int y = (int)(r % Ho);
    long long pl = r / Ho;
    out[i] = in[(pl * Hi + y / f) * Wi + x / f] * scale;
  }
}

// out[k] = scale * sum_{n,hw} g[n,k,hw]  (to_rgb bias gradient); grid (K, PS_CHUNKS) partials, then a final pass
constexpr int PS_CHUNKS = 64;
__global__ void plane_sum_kernel(const float* __restrict__ g, float* __restrict__ part, int N, int K, int HW) {
  __shared__ float sm[32];
  const int k = blockIdx.x, ch = blockIdx.y;
  const long long total = (long long)N * HW;
  float s = 0.f;
  for (long long i = (long long)ch * blockDim.x + threadIdx.x; i < total; i += (long long)PS_CHUNKS * blockDim.x) {
    long long n = i / HW;
    s += g[((size_t)n * K + k) * HW + (i - n * HW)];
  }
  s = sg_block_sum(s, sm);
  if (threadIdx.x == 0) part[k * PS_CHUNKS + ch] = s;
}
__global__ void plane_sum_final_kernel(const float* __restrict__ part, float* __restrict__ out, int K, float scale,
                                       int accumulate) {
  int k = threadIdx.x;
  if (k >= K) return;
  float s = 0.f;
  for (int c = 0; c < PS_CHUNKS; ++c) s += part[k * PS_CHUNKS + c];
  out[k] = (accumulate ? out[k] : 0.f) + s * scale;
}

// ---------------------------------------------------------------------------------------------------------
// dlatent assembly: style mixing + truncation.   reference GAN.py:281-293, CustomLayers.py:319-323
//   src = l < mix_cutoff ? w1[b] : w2[b];   out[b,l] = l < trunc_cutoff ? avg + psi*(src-avg) : src
// ---------------------------------------------------------------------------------------------------------
__global__ void dlatents_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                const float* __restrict__ avg, float* __restrict__ out, int B, int L, int D,
                                int mix_cutoff, float psi, int trunc_cutoff, const int* __restrict__ cutoff_dev = nullptr) {
  if (cutoff_dev) mix_cutoff = *cutoff_dev;
  const long long total = (long long)B * L * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int d = (int)(i % D);
    long long r = i / D;
    int l = (int)(r % L);
    int b = (int)(r / L);
    float s = (w2 == nullptr || l < mix_cutoff) ? w1[(size_t)b * D + d] : w2[(size_t)b * D + d];
    if (avg != nullptr && l < trunc_cutoff) { float a = avg[d]; s = a + psi * (s - a); }
    out[i] = s;
  }
}

__global__ void dlatents_bwd_kernel(const float* __restrict__ g, float* __restrict__ gw1, float* __restrict__ gw2,
                                    int B, int L, int D, int mix_cutoff, float psi, int trunc_cutoff, int has_trunc,
                                    const int* __restrict__ cutoff_dev = nullptr) {
  if (cutoff_dev) mix_cutoff = *cutoff_dev;
  const long long total = (long long)B * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int d = (int)(i % D);
    int b = (int)(i / D);
    float s1 = 0.f, s2 = 0.f;
    for (int l = 0; l < L; ++l) {
      float v = g[((size_t)b * L + l) * D + d] * ((has_trunc && l < trunc_cutoff) ? psi : 1.f);
      if (gw2 == nullptr || l < mix_cutoff) s1 += v; else s2 += v;
    }
    gw1[i] = s1;
    if (gw2) gw2[i] = s2;
  }
}

// ---------------------------------------------------------------------------------------------------------
// NHWC 2x nearest up-sampling (times scale) and 2x2 box-sum pooling (times scale, + optional bias and LeakyReLU):
// an adjoint pair.  Used to put the reference's unfused up/down convolutions (CustomLayers.py:153-154, 166-176)
// on the tensor-core 3x3 kernel: nearest-up -> conv3x3, conv3x3 -> 2x2 box.
// ---------------------------------------------------------------------------------------------------------
template <class T>
__global__ void up2_nhwc_kernel(const T* __restrict__ x, T* __restrict__ y, float scale, int N, int H, int W, int C) {
  using V = SgVec<T>;
  const int CV = C / V::N;
  const int Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)N * Ho * Wo * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    long long p = i / CV;
    int ox = (int)(p % Wo);
    long long r = p / Wo;
    int oy = (int)(r % Ho);
    int n = (int)(r / Ho);
    V v, o;
    v.load(x + (((size_t)n * H + (oy >> 1)) * W + (ox >> 1)) * C + cv * V::N);
    if (scale == 1.f) o = v;
    else {
#pragma unroll
      for (int k = 0; k < V::N; ++k) o.set(k, v.get(k) * scale);
    }
    o.store(y + (size_t)p * C + cv * V::N);
  }
}

template <class T>
__global__ void pool2_nhwc_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ bias,
                                  float bias_mul, int act, float scale, int N, int Ho, int Wo, int C) {
  using V = SgVec<T>;
  const int CV = C / V::N;
  const int H = 2 * Ho, W = 2 * Wo;
  const long long total = (long long)N * Ho * Wo * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    long long p = i / CV;
    int ox = (int)(p % Wo);
    long long r = p / Wo;
    int oy = (int)(r % Ho);
    int n = (int)(r / Ho);
    const T* src = x + (((size_t)n * H + 2 * oy) * W + 2 * ox) * C + cv * V::N;
    V a, b, c, d, o;
    a.load(src); b.load(src + C); c.load(src + (size_t)W * C); d.load(src + (size_t)W * C + C);
#pragma unroll
    for (int k = 0; k < V::N; ++k) {
      float v = (a.get(k) + b.get(k) + c.get(k) + d.get(k)) * scale;
      if (bias) v = fmaf(bias[cv * V::N + k], bias_mul, v);
      if (act) v = sg_lrelu(v);
      o.set(k, v);
    }
    o.store(y + (size_t)p * C + cv * V::N);
  }
}

// dst[n, y, x, :] = (y < hs && x < ws) ? src[n, y, x, :] : 0 -- zero-pad (dst larger) or crop (dst smaller) of an NHWC
// map.  Used to run the 4x4 / 8x8 layers on the 8x16-pixel tiles of the tensor-core kernels.
template <class T>
__global__ void pad_crop_nhwc_kernel(const T* __restrict__ src, T* __restrict__ dst, int N, int hs, int ws, int hd,
                                     int wd, int C) {
  using V = SgVec<T>;
  const int CV = C / V::N;
  const long long total = (long long)N * hd * wd * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(i % CV);
    long long p = i / CV;
    int x = (int)(p % wd);
    long long r = p / wd;
    int y = (int)(r % hd);
    int n = (int)(r / hd);
    V v;
    if (y < hs && x < ws) v.load(src + (((size_t)n * hs + y) * ws + x) * C + cv * V::N);
    else v.zero();
    v.store(dst + (size_t)p * C + cv * V::N);
  }
}

// y = x + bias[c]*bias_mul + noise_w[c]*noise[n,p]   (stand-alone NoiseLayer / bias add; CustomLayers.py:191-200,178)
template <class T>
__global__ void bias_noise_kernel(const T* __restrict__ x, const float* __restrict__ bias, float bias_mul,
                                  const float* __restrict__ noise, const float* __restrict__ noise_w,
                                  T* __restrict__ y, long long P, int C, int act) {
  const long long total = P * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long p = i / C;
    float v = sg_to_f(x[i]);
    if (bias) v = fmaf(bias[c], bias_mul, v);
    if (noise) v = fmaf(noise_w[c], noise[p], v);
    if (act) v = sg_lrelu(v);
    y[i] = sg_from_f<T>(v);
  }
}

}  // namespace

SG_EXPORT int sg_bias_noise(const void* x, const float* bias, float bias_mul, const float* noise,
                            const float* noise_w, int act, void* y, int dtype, long long P, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_DISPATCH(dtype, {
    bias_noise_kernel<T><<<sg_grid_1d(P * C, 256, 148 * 32), 256, 0, st>>>((const T*)x, bias, bias_mul, noise, noise_w, (T*)y, P, C, act);
  });
  return sg_check_launch("sg_bias_noise");
}

SG_EXPORT int sg_blur3(const void* x, void* y, const float* bias, float bias_mul, int dtype, int N, int H, int W,
                       int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (g_sg_fast_paths) {
    int rc = sgf_blur3(x, y, bias, bias_mul, dtype, N, H, W, C, stream);
    if (rc != SG_FAST_UNSUPPORTED) { if (rc) sg_set_error("sg_blur3/v2: launch failed"); return rc; }
  }
  SG_DISPATCH(dtype, {
    if (C % SgVec<T>::N == 0) {
      long long total = (long long)N * ((H + BLUR_STRIP - 1) / BLUR_STRIP) * W * (C / SgVec<T>::N);
      blur3_kernel<T><<<sg_grid_1d(total, 128, 148 * 32), 128, 0, st>>>((const T*)x, (T*)y, bias, bias_mul, N, H, W, C);
    } else {
      long long total = (long long)N * H * W * C;
      blur3_scalar_kernel<T><<<sg_grid_1d(total, 256, 148 * 32), 256, 0, st>>>((const T*)x, (T*)y, bias, bias_mul, N, H, W, C);
    }
  });
  return sg_check_launch("sg_blur3");
}

SG_EXPORT int sg_axpby(const void* x, const void* y, void* out, float a, float b, long long n, int dtype,
                       void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_DISPATCH(dtype, {
    axpby_kernel<T><<<sg_grid_1d(n / SgVec<T>::N + 1, 256, 148 * 32), 256, 0, st>>>((const T*)x, (const T*)y, (T*)out, a, b, n);
  });
  return sg_check_launch("sg_axpby");
}

SG_EXPORT int sg_axpby_dev(const void* x, const void* y, void* out, const float* ab, long long n, int dtype,
                           void* stream) {
  SG_REQUIRE(ab != nullptr, "sg_axpby_dev: coefficient pointer missing");
  cudaStream_t st = (cudaStream_t)stream;
  SG_DISPATCH(dtype, {
    axpby_kernel<T><<<sg_grid_1d(n / SgVec<T>::N + 1, 256, 148 * 32), 256, 0, st>>>((const T*)x, (const T*)y, (T*)out, 0.f, 0.f, n, ab);
  });
  return sg_check_launch("sg_axpby_dev");
}

SG_EXPORT int sg_lrelu_bwd(const void* g, const void* ref, void* gx, long long n, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_DISPATCH(dtype, {
    lrelu_bwd_kernel<T><<<sg_grid_1d(n / SgVec<T>::N + 1, 256, 148 * 32), 256, 0, st>>>((const T*)g, (const T*)ref, (T*)gx, n);
  });
  return sg_check_launch("sg_lrelu_bwd");
}

// gx = g * lrelu'(ref) over [P][C] (NHWC) and gb[c] = scale * sum_p gx[p][c] in the same pass (two-stage, deterministic).
// ws: sg_channel_sum_blocks(P) * C floats.  1 = supported (C a multiple of the 16-byte vector, <= 2048 channels).
SG_EXPORT int sg_lrelu_bwd_bias_supported(int dtype, long long P, int C) {
  const int VN = dtype == SG_BF16 ? 8 : 4;
  if (C % VN != 0 || P < 1 || P >= (1ll << 31)) return 0;
  const int CV = C / VN;
  return CV <= 256 && (size_t)(256 / CV) * C * sizeof(float) <= 48 * 1024;
}
SG_EXPORT int sg_lrelu_bwd_bias(const void* g, const void* ref, void* gx, int dtype, long long P, int C, float* ws,
                                float* gb, float scale, void* stream) {
  SG_REQUIRE(sg_lrelu_bwd_bias_supported(dtype, P, C), "sg_lrelu_bwd_bias: shape not supported");
  SG_REQUIRE(ws != nullptr && gb != nullptr, "sg_lrelu_bwd_bias: workspace / output missing");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = sg_channel_sum_blocks(P);
  const int pix_per_blk = (int)((P + nblk - 1) / nblk);
  SG_DISPATCH(dtype, {
    const int CV = C / SgVec<T>::N;
    const size_t smv = (size_t)(256 / CV) * C * sizeof(float);
    lrelu_bwd_bias_kernel<T><<<nblk, 256, smv, st>>>((const T*)g, (const T*)ref, (T*)gx, P, C, ws, pix_per_blk);
  });
  int rc = sg_check_launch("sg_lrelu_bwd_bias");
  if (rc) return rc;
  partial_reduce_kernel<<<sg_cdiv(C, 4), 128, 0, st>>>(ws, nblk, C, gb, scale, 0);
  return sg_check_launch("sg_lrelu_bwd_bias/reduce");
}

SG_EXPORT int sg_cast(const void* x, int dtype_in, void* y, int dtype_out, long long n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int grid = sg_grid_1d(n, 256, 148 * 32);
  if (dtype_in == SG_F32 && dtype_out == SG_BF16)
    cast_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)x, (__nv_bfloat16*)y, n);
  else if (dtype_in == SG_BF16 && dtype_out == SG_F32)
    cast_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (float*)y, n);
  else if (dtype_in == SG_F32 && dtype_out == SG_F32)
    cast_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, n);
  else if (dtype_in == SG_BF16 && dtype_out == SG_BF16)
    cast_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n);
  else { sg_set_error("sg_cast: bad dtype"); return 1001; }
  return sg_check_launch("sg_cast");
}

SG_EXPORT int sg_channel_sum_blocks(long long P) {
  long long b = (P + 255) / 256;
  if (b > 148 * 4) b = 148 * 4;
  if (b < 1) b = 1;
  return (int)b;
}

SG_EXPORT int sg_channel_sum(const void* x, int dtype, const float* wt, int J, int ones, long long P, int HW, int C,
                             float* part_ws, float* out, float scale, int accumulate, void* stream) {
  int rows = J + (ones ? 1 : 0);
  SG_REQUIRE(rows >= 1 && rows <= CS_ROWS && J >= 0, "sg_channel_sum: rows must be 1..4");
  SG_REQUIRE(J == 0 || wt != nullptr, "sg_channel_sum: weights missing");
  SG_REQUIRE(P < (1ll << 31), "sg_channel_sum: too many pixels");
  cudaStream_t st = (cudaStream_t)stream;
  if (g_sg_fast_paths && dtype == SG_BF16 && J == 1 && ones && !accumulate) {
    // noise-weight + bias gradient: one per-pixel weight row plus the plain sum (same workspace layout and blocks)
    int rc = sgf_channel_sum_w1(x, wt, P, C, part_ws, out, scale, stream);
    if (rc != SG_FAST_UNSUPPORTED) { if (rc) sg_set_error("sg_channel_sum/w1: launch failed"); return rc; }
  }
  if (g_sg_fast_paths && dtype == SG_BF16 && J == 3 && !accumulate) {
    // rgb weight gradients: three per-pixel weight rows (+ the plain sum for from_rgb's bias)
    int rc = sgf_channel_sum_w3(x, wt, ones, P, HW, C, part_ws, out, scale, stream);
    if (rc != SG_FAST_UNSUPPORTED) { if (rc) sg_set_error("sg_channel_sum/w3: launch failed"); return rc; }
  }
  int nblk = sg_channel_sum_blocks(P);
  int pix_per_blk = (int)((P + nblk - 1) / nblk);
  int bx = C >= 128 ? 128 : (C >= 64 ? 64 : 32);
  int R = 256 / bx;
  size_t smem = (size_t)rows * R * C * sizeof(float);
  while (smem > 48 * 1024 && R > 1) { R >>= 1; smem = (size_t)rows * R * C * sizeof(float); }
  SG_REQUIRE(smem <= 48 * 1024, "sg_channel_sum: C too large");
  dim3 block(bx, R);
  SG_DISPATCH(dtype, {
    const int VN = SgVec<T>::N;
    const int CV = C / VN;
    if (C % VN == 0 && CV <= 256 && (size_t)rows * (256 / CV) * C * sizeof(float) <= 48 * 1024) {
      size_t smv = (size_t)rows * (256 / CV) * C * sizeof(float);
      if (J == 0)
        channel_sum_ones_kernel<T><<<nblk, 256, smv, st>>>((const T*)x, P, C, part_ws, pix_per_blk);
      else
        channel_sum_vec_kernel<T><<<nblk, 256, smv, st>>>((const T*)x, wt, J, ones, P, HW, C, part_ws, pix_per_blk);
    } else {
      channel_sum_kernel<T><<<nblk, block, smem, st>>>((const T*)x, wt, J, ones, P, HW, C, part_ws, pix_per_blk);
    }
  });
  int rc = sg_check_launch("sg_channel_sum");
  if (rc) return rc;
  int M = rows * C;
  partial_reduce_kernel<<<sg_cdiv(M, 4), 128, 0, st>>>(part_ws, nblk, M, out, scale, accumulate);
  return sg_check_launch("sg_channel_sum/reduce");
}

SG_EXPORT int sg_rgb_in(const float* img, const float* W, const float* bias, float scale, float bias_mul, int act,
                        void* y, int dtype, int N, int HW, int K, int C, void* stream) {
  SG_REQUIRE(K >= 1 && K <= 4, "sg_rgb_in: K must be 1..4");
  cudaStream_t st = (cudaStream_t)stream;
  if (g_sg_fast_paths && dtype == SG_BF16 && K == 3) {
    int rc = sgf_rgb_in3(img, W, bias, scale, bias_mul, act, y, N, HW, C, stream);
    if (rc != SG_FAST_UNSUPPORTED) { if (rc) sg_set_error("sg_rgb_in/k3: launch failed"); return rc; }
  }
  long long total = (long long)N * HW * C;
  SG_DISPATCH(dtype, {
    if (C % SgVec<T>::N == 0 && 256 % (C / SgVec<T>::N) == 0 && total / SgVec<T>::N < (1ll << 30))   // CV | 256 -> fixed cv per thread
      rgb_in_vec_kernel<T><<<sg_grid_1d(total / SgVec<T>::N, 256, 148 * 32), 256, 0, st>>>(img, W, bias, scale, bias_mul, act, (T*)y, N, HW, K, C);
    else
      rgb_in_kernel<T><<<sg_grid_1d(total, 256, 148 * 32), 256, 0, st>>>(img, W, bias, scale, bias_mul, act, (T*)y, N, HW, K, C);
  });
  return sg_check_launch("sg_rgb_in");
}

SG_EXPORT int sg_rgb_in_dgrad(const void* g, int dtype, const float* W, float scale, float* gimg, int N, int HW,
                              int K, int C, void* stream) {
  SG_REQUIRE(K >= 1 && K <= 4, "sg_rgb_in_dgrad: K must be 1..4");
  cudaStream_t st = (cudaStream_t)stream;
  long long P = (long long)N * HW;
  SG_DISPATCH(dtype, {
    SG_REQUIRE(C % SgVec<T>::N == 0, "sg_rgb_in_dgrad: C must be a multiple of the 16-byte vector width");
    int G = 1;
    while (G < 32 && G < C / SgVec<T>::N) G <<= 1;
    if (G < 4) G = 4;                       // K <= 4 writer lanes
    SG_REQUIRE(P < (1ll << 30), "sg_rgb_in_dgrad: too many pixels");
    rgb_in_dgrad_kernel<T><<<sg_grid_1d((P * G + 3) / 4, 256, 148 * 8), 256, 0, st>>>((const T*)g, W, scale, gimg, N, HW, K, C, G);
  });
  return sg_check_launch("sg_rgb_in_dgrad");
}

SG_EXPORT int sg_rgb_out(const void* x, int dtype, const float* W, const float* bias, float scale, float bias_mul,
                         float alpha, float beta, float* out, int N, int h, int w, int C, int K, int up, void* stream) {
  SG_REQUIRE(K >= 1 && K <= 4 && up >= 1 && up <= 2, "sg_rgb_out: K must be 1..4, up 1..2");
  cudaStream_t st = (cudaStream_t)stream;
  if (g_sg_fast_paths && dtype == SG_BF16 && K == 3) {
    int rc = sgf_rgb_out3(x, W, bias, scale, bias_mul, alpha, beta, out, N, h, w, C, up, stream);
    if (rc != SG_FAST_UNSUPPORTED) { if (rc) sg_set_error("sg_rgb_out/k3: launch failed"); return rc; }
  }
  long long P = (long long)N * h * w;
  SG_DISPATCH(dtype, {
    SG_REQUIRE(C % SgVec<T>::N == 0, "sg_rgb_out: C must be a multiple of the 16-byte vector width");
    int G = 1;
    while (G < 32 && G < C / SgVec<T>::N) G <<= 1;
    SG_REQUIRE(P < (1ll << 30), "sg_rgb_out: too many pixels");
    rgb_out_kernel<T><<<sg_grid_1d((P * G + 3) / 4, 256, 148 * 8), 256, 0, st>>>((const T*)x, W, bias, scale, bias_mul, alpha, beta, out, N, h, w, C, K, up, G);
  });
  return sg_check_launch("sg_rgb_out");
}

SG_EXPORT int sg_rgb_out_dgrad(const float* g, const float* W, float coef, void* gx, int dtype, int N, int h, int w,
                               int C, int K, int up, void* stream) {
  SG_REQUIRE(K >= 1 && K <= 4 && up >= 1 && up <= 2, "sg_rgb_out_dgrad: K must be 1..4, up 1..2");
  cudaStream_t st = (cudaStream_t)stream;
  SG_DISPATCH(dtype, {
    SG_REQUIRE(C % SgVec<T>::N == 0, "sg_rgb_out_dgrad: C must be a multiple of the 16-byte vector width");
    long long total = (long long)N * h * w * (C / SgVec<T>::N);
    if (256 % (C / SgVec<T>::N) == 0 && total < (1ll << 30))
      rgb_out_dgrad_vec_kernel<T><<<sg_grid_1d(total, 256, 148 * 32), 256, 0, st>>>(g, W, coef, (T*)gx, N, h, w, C, K, up);
    else
      rgb_out_dgrad_kernel<T><<<sg_grid_1d(total, 256, 148 * 32), 256, 0, st>>>(g, W, coef, (T*)gx, N, h, w, C, K, up);
  });
  return sg_check_launch("sg_rgb_out_dgrad");
}

SG_EXPORT int sg_pool_nchw(const float* in, float* out, long long planes, int Ho, int Wo, int f, float scale,
                           void* stream) {
  long long total = planes * Ho * Wo;
  pool_nchw_kernel<<<sg_grid_1d(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(in, out, planes, Ho, Wo, f, scale);
  return sg_check_launch("sg_pool_nchw");
}

SG_EXPORT int sg_upsample_nchw(const float* in, float* out, long long planes, int Hi, int Wi, int f, float scale,
                               void* stream) {
  long long total = planes * Hi * f * Wi * f;
  upsample_nchw_kernel<<<sg_grid_1d(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(in, out, planes, Hi, Wi, f, scale);
  return sg_check_launch("sg_upsample_nchw");
}

// ws: K*64 floats
SG_EXPORT int sg_plane_sum(const float* g, float* out, float* ws, int N, int K, int HW, float scale, int accumulate,
                           void* stream) {
  SG_REQUIRE(K >= 1 && K <= 32, "sg_plane_sum: K must be 1..32");
  cudaStream_t st = (cudaStream_t)stream;
  plane_sum_kernel<<<dim3(K, PS_CHUNKS), 256, 0, st>>>(g, ws, N, K, HW);
  plane_sum_final_kernel<<<1, 32, 0, st>>>(ws, out, K, scale, accumulate);
  return sg_check_launch("sg_plane_sum");
}

SG_EXPORT int sg_dlatents(const float* w1, const float* w2, const float* avg, float* out, int B, int L, int D,
                          int mix_cutoff, float psi, int trunc_cutoff, void* stream) {
  long long total = (long long)B * L * D;
  dlatents_kernel<<<sg_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(w1, w2, avg, out, B, L, D, mix_cutoff, psi, trunc_cutoff);
  return sg_check_launch("sg_dlatents");
}

SG_EXPORT int sg_dlatents_bwd(const float* g, float* gw1, float* gw2, int B, int L, int D, int mix_cutoff, float psi,
                              int trunc_cutoff, int has_trunc, void* stream) {
  long long total = (long long)B * D;
  dlatents_bwd_kernel<<<sg_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(g, gw1, gw2, B, L, D, mix_cutoff, psi, trunc_cutoff, has_trunc);
  return sg_check_launch("sg_dlatents_bwd");
}

// the style-mixing cut-off read from device memory (one int), for CUDA-graph replay
SG_EXPORT int sg_dlatents_dev(const float* w1, const float* w2, const float* avg, float* out, int B, int L, int D,
                              const int* mix_cutoff, float psi, int trunc_cutoff, void* stream) {
  SG_REQUIRE(mix_cutoff != nullptr, "sg_dlatents_dev: cut-off pointer missing");
  long long total = (long long)B * L * D;
  dlatents_kernel<<<sg_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(w1, w2, avg, out, B, L, D, 0, psi, trunc_cutoff, mix_cutoff);
  return sg_check_launch("sg_dlatents_dev");
}
SG_EXPORT int sg_dlatents_bwd_dev(const float* g, float* gw1, float* gw2, int B, int L, int D, const int* mix_cutoff,
                                  float psi, int trunc_cutoff, int has_trunc, void* stream) {
  SG_REQUIRE(mix_cutoff != nullptr, "sg_dlatents_bwd_dev: cut-off pointer missing");
  long long total = (long long)B * D;
  dlatents_bwd_kernel<<<sg_grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(g, gw1, gw2, B, L, D, 0, psi, trunc_cutoff, has_trunc, mix_cutoff);
  return sg_check_launch("sg_dlatents_bwd_dev");
}

SG_EXPORT int sg_up2_nhwc(const void* x, void* y, float scale, int dtype, int N, int H, int W, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_DISPATCH(dtype, {
    SG_REQUIRE(C % SgVec<T>::N == 0, "sg_up2_nhwc: C must be a multiple of the 16-byte vector width");
    long long total = (long long)N * 4 * H * W * (C / SgVec<T>::N);
    up2_nhwc_kernel<T><<<sg_grid_1d(total, 256, 148 * 32), 256, 0, st>>>((const T*)x, (T*)y, scale, N, H, W, C);
  });
  return sg_check_launch("sg_up2_nhwc");
}

SG_EXPORT int sg_pool2_nhwc(const void* x, void* y, const float* bias, float bias_mul, int act, float scale,
                            int dtype, int N, int Ho, int Wo, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_DISPATCH(dtype, {
    SG_REQUIRE(C % SgVec<T>::N == 0, "sg_pool2_nhwc: C must be a multiple of the 16-byte vector width");
    long long total = (long long)N * Ho * Wo * (C / SgVec<T>::N);
    pool2_nhwc_kernel<T><<<sg_grid_1d(total, 256, 148 * 32), 256, 0, st>>>((const T*)x, (T*)y, bias, bias_mul, act, scale, N, Ho, Wo, C);
  });
  return sg_check_launch("sg_pool2_nhwc");
}

SG_EXPORT int sg_pad_crop_nhwc(const void* src, void* dst, int dtype, int N, int hs, int ws, int hd, int wd, int C,
                               void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_DISPATCH(dtype, {
    SG_REQUIRE(C % SgVec<T>::N == 0, "sg_pad_crop_nhwc: C must be a multiple of the 16-byte vector width");
    long long total = (long long)N * hd * wd * (C / SgVec<T>::N);
    pad_crop_nhwc_kernel<T><<<sg_grid_1d(total, 256, 148 * 16), 256, 0, st>>>((const T*)src, (T*)dst, N, hs, ws, hd, wd, C);
  });
  return sg_check_launch("sg_pad_crop_nhwc");
}
