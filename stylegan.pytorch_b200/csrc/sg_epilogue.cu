// Synthesis-layer epilogue (reference CustomLayers.py:219-248: noise -> LeakyReLU -> InstanceNorm -> StyleMod)
// as two forward passes and two backward passes over NHWC activations:
//   fwd 1  z = lrelu(y + b[c] + nw[c]*noise[n,p]);  S[n,c] += (z, z^2)            (sg_epi_act_stats)
//   fwd 2  out = (z - mean[n,c])*a[n,c] + s1[n,c]   with a = rstd*(1+s0)          (sg_epi_finalize + sg_affine_nc)
//   bwd 1  A[n,c] = sum_p g,  B[n,c] = sum_p g*xhat                               (sg_epi_bwd_reduce)
//   bwd 2  g_pre = (c1*g + k*(z-mean) + c0) * lrelu'(z)                           (sg_epi_bwd_finalize + _apply)
// Numerics: the statistics are sums of (z - K) with the pivot K[n,c] = z at the image's first pixel, so the one-pass
// variance does not cancel even when mean^2 >> var (the 4x4 constant-input layer: z = 2 + small noise); block
// partials are fp32 (one record per block, no atomics), summed in block order in fp64 by a second small kernel: bitwise
// reproducible.  All affine forms are written around (z - mean) for the same reason.
#include "sg_common.cuh"
#include "../../include/sgb200.h"

namespace {

constexpr int EPI_ROWS = 8;  // pixel rows per block (blockDim.y)

// Two-value reduction over the pixels of one image: grid (chunks, N), block (CX, EPI_ROWS).
// MODE 0: v = lrelu(y + b + nw*noise) -> write z, accumulate (z, z^2)
// MODE 1: accumulate (g, g*xhat), xhat = (z-mean)*rstd
template <class T, int MODE>
__global__ void epi_reduce_kernel(const T* __restrict__ in0, const T* __restrict__ in1, T* __restrict__ zout,
                                  const float* __restrict__ bias, float bias_mul, const float* __restrict__ noise,
                                  const float* __restrict__ noise_w, const float* __restrict__ mr,
                                  double* __restrict__ sums, float* __restrict__ ws, int HW, int C, int pix_per_blk,
                                  int act) {
  extern __shared__ float sm[];  // [2][EPI_ROWS][C]
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_blk;
  int p1 = p0 + pix_per_blk;
  if (p1 > HW) p1 = HW;
  const size_t base = (size_t)n * HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    float b = 0.f, nw = 0.f, mean = 0.f, rstd = 0.f, pivot = 0.f;
    if (MODE == 0) {
      b = bias ? bias[c] * bias_mul : 0.f;
      nw = (noise && noise_w) ? noise_w[c] : 0.f;
      // pivot = z at pixel 0 of this image (identical in every CTA of the image)
      float v0 = sg_to_f(in0[base * C + c]) + b;
      if (noise) v0 = fmaf(nw, noise[base], v0);
      if (act) v0 = sg_lrelu(v0);
      pivot = sg_to_f(sg_from_f<T>(v0));
      if (blockIdx.x == 0 && threadIdx.y == 0) sums[((size_t)n * C + c) * 3 + 2] = (double)pivot;
    } else {
      mean = mr[((size_t)n * C + c) * 2 + 0];
      rstd = mr[((size_t)n * C + c) * 2 + 1];
    }
    for (int p = p0 + threadIdx.y; p < p1; p += EPI_ROWS) {
      size_t idx = (base + p) * C + c;
      if (MODE == 0) {
        float v = sg_to_f(in0[idx]) + b;
        if (noise) v = fmaf(nw, noise[base + p], v);
        if (act) v = sg_lrelu(v);
        T zt = sg_from_f<T>(v);
        zout[idx] = zt;
        float zr = sg_to_f(zt) - pivot;
        s0 += zr;
        s1 = fmaf(zr, zr, s1);
      } else {
        float g = sg_to_f(in0[idx]);
        float xh = (sg_to_f(in1[idx]) - mean) * rstd;
        s0 += g;
        s1 = fmaf(g, xh, s1);
      }
    }
    sm[(0 * EPI_ROWS + threadIdx.y) * C + c] = s0;
    sm[(1 * EPI_ROWS + threadIdx.y) * C + c] = s1;
  }
  __syncthreads();
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
  for (int i = tid; i < 2 * C; i += nthr) {
    int j = i / C, c = i % C;
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < EPI_ROWS; ++r) s += sm[(j * EPI_ROWS + r) * C + c];
    ws[(((size_t)n * gridDim.x + blockIdx.x) * 2 + j) * C + c] = s;
  }
}

// Vectorised variant: a thread owns one 16-byte channel vector and strides over the pixels of its chunk.
template <class T, int MODE>
__global__ void __launch_bounds__(256) epi_reduce_vec_kernel(const T* __restrict__ in0, const T* __restrict__ in1,
                                                             T* __restrict__ zout, const float* __restrict__ bias,
                                                             float bias_mul, const float* __restrict__ noise,
                                                             const float* __restrict__ noise_w,
                                                             const float* __restrict__ mr, double* __restrict__ sums,
                                                             float* __restrict__ ws, int HW, int C, int pix_per_blk,
                                                             int act) {
  using V = SgVec<T>;
  extern __shared__ float sm[];  // [2][PR][C]
  const int CV = C / V::N;
  const int PR = blockDim.x / CV;
  const int cv = threadIdx.x % CV, pr = threadIdx.x / CV;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_blk;
  int p1 = p0 + pix_per_blk;
  if (p1 > HW) p1 = HW;
  const size_t base = (size_t)n * HW;
  if (pr < PR) {
    float s0[V::N], s1[V::N], b[V::N], nw[V::N], a0[V::N], a1[V::N];   // a0/a1: pivot (MODE 0) or mean/rstd (MODE 1)
#pragma unroll
    for (int k = 0; k < V::N; ++k) {
      const int c = cv * V::N + k;
      s0[k] = s1[k] = 0.f;
      if (MODE == 0) {
        b[k] = bias ? bias[c] * bias_mul : 0.f;
        nw[k] = (noise && noise_w) ? noise_w[c] : 0.f;
      } else {
        a0[k] = mr[((size_t)n * C + c) * 2 + 0];
        a1[k] = mr[((size_t)n * C + c) * 2 + 1];
      }
    }
    if (MODE == 0) {
      V v0;
      v0.load(in0 + base * C + cv * V::N);
      const float nz0 = noise ? noise[base] : 0.f;
#pragma unroll
      for (int k = 0; k < V::N; ++k) {
        float v = v0.get(k) + b[k];
        if (noise) v = fmaf(nw[k], nz0, v);
        if (act) v = sg_lrelu(v);
        a0[k] = sg_to_f(sg_from_f<T>(v));
        if (blockIdx.x == 0 && pr == 0) sums[((size_t)n * C + cv * V::N + k) * 3 + 2] = (double)a0[k];
      }
    }
    constexpr int U = 4;                          // pixels in flight per thread
    for (int pb = p0 + pr; pb < p1; pb += U * PR) {
      V va[U], vb[U];
      float nz[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int p = pb + u * PR;
        ok[u] = p < p1;
        nz[u] = 0.f;
        if (ok[u]) {
          const size_t idx = (base + p) * C + cv * V::N;
          va[u].load(in0 + idx);
          if (MODE == 0) { if (noise) nz[u] = noise[base + p]; }
          else vb[u].load(in1 + idx);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        const size_t idx = (base + pb + u * PR) * C + cv * V::N;
        if (MODE == 0) {
          V vo;
#pragma unroll
          for (int k = 0; k < V::N; ++k) {
            float v = va[u].get(k) + b[k];
            if (noise) v = fmaf(nw[k], nz[u], v);
            if (act) v = sg_lrelu(v);
            vo.set(k, v);
          }
          vo.store(zout + idx);
#pragma unroll
          for (int k = 0; k < V::N; ++k) {
            float zr = vo.get(k) - a0[k];
            s0[k] += zr;
            s1[k] = fmaf(zr, zr, s1[k]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < V::N; ++k) {
            float g = va[u].get(k);
            float xh = (vb[u].get(k) - a0[k]) * a1[k];
            s0[k] += g;
            s1[k] = fmaf(g, xh, s1[k]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < V::N; ++k) {
      sm[((size_t)0 * PR + pr) * C + cv * V::N + k] = s0[k];
      sm[((size_t)1 * PR + pr) * C + cv * V::N + k] = s1[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    int j = i / C, c = i % C;
    float s = 0.f;
    for (int r = 0; r < PR; ++r) s += sm[((size_t)j * PR + r) * C + c];
    ws[(((size_t)n * gridDim.x + blockIdx.x) * 2 + j) * C + c] = s;
  }
}

// sums[n,c,j] = sum over the image's blocks, in block order, of the fp32 partials (fp64 accumulation)
__global__ void epi_sum_parts_kernel(const float* __restrict__ ws, int chunks, double* __restrict__ sums, int N, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C * 2) return;
  const int c = i % C, j = (i / C) % 2, n = i / (2 * C);
  double s = 0.0;
  for (int k = 0; k < chunks; ++k) s += (double)ws[(((size_t)n * chunks + k) * 2 + j) * C + c];
  sums[((size_t)n * C + c) * 3 + j] = s;
}

// per-(n,c): mean/rstd and the fused affine (a, d).  style = [N, 2C] (s0 | s1); style may be null (plain IN).
__global__ void epi_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ style,
                                    float* __restrict__ mr, float* __restrict__ ab, int N, int C, int HW, float eps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  int n = i / C, c = i % C;
  double m = sums[(size_t)i * 3] / HW;                       // mean of (z - pivot)
  double var = sums[(size_t)i * 3 + 1] / HW - m * m;
  if (var < 0.0) var = 0.0;
  float rstd = (float)(1.0 / sqrt(var + (double)eps));
  float mean = (float)(m + sums[(size_t)i * 3 + 2]);
  mr[(size_t)i * 2] = mean;
  mr[(size_t)i * 2 + 1] = rstd;
  float s0 = style ? style[(size_t)n * 2 * C + c] : 0.f;
  float s1 = style ? style[(size_t)n * 2 * C + C + c] : 0.f;
  ab[(size_t)i * 3] = mean;
  ab[(size_t)i * 3 + 1] = rstd * (1.f + s0);
  ab[(size_t)i * 3 + 2] = s1;
}

// out[n,p,c] = (z - mab[n,c,0]) * mab[n,c,1] + mab[n,c,2]
// grid (pixel chunks, N), blockDim a multiple of CV: a thread keeps one channel vector of one image for its whole
// loop, so its 3*VN coefficients are loaded once into registers (a per-pixel gather of them from L1 -- 8 distinct
// lines per warp instruction -- held these kernels at 1.6-4 TB/s).
template <class T>
__global__ void __launch_bounds__(256) affine_nc_kernel(const T* __restrict__ z, const float* __restrict__ ab,
                                                        T* __restrict__ out, int HW, int C, int pix_per_blk) {
  using V = SgVec<T>;
  const int CV = C / V::N;
  const int n = blockIdx.y;
  const int cv = threadIdx.x % CV, pr = threadIdx.x / CV, PR = blockDim.x / CV;
  const int p0 = blockIdx.x * pix_per_blk;
  int p1 = p0 + pix_per_blk;
  if (p1 > HW) p1 = HW;
  float q[3 * V::N];
#pragma unroll
  for (int j = 0; j < 3 * V::N; ++j) q[j] = ab[((size_t)n * C + cv * V::N) * 3 + j];
  const size_t base = ((size_t)n * HW) * C + cv * V::N;
  constexpr int U = 4;
  int p = p0 + pr;
  for (; p + (U - 1) * PR < p1; p += U * PR) {
    V v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u].load(z + base + (size_t)(p + u * PR) * C);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      V o;
#pragma unroll
      for (int k = 0; k < V::N; ++k) o.set(k, fmaf(v[u].get(k) - q[3 * k], q[3 * k + 1], q[3 * k + 2]));
      o.store(out + base + (size_t)(p + u * PR) * C);
    }
  }
  for (; p < p1; p += PR) {
    V v, o;
    v.load(z + base + (size_t)p * C);
#pragma unroll
    for (int k = 0; k < V::N; ++k) o.set(k, fmaf(v.get(k) - q[3 * k], q[3 * k + 1], q[3 * k + 2]));
    o.store(out + base + (size_t)p * C);
  }
}
template <class T>
__global__ void affine_nc_scalar_kernel(const T* __restrict__ z, const float* __restrict__ ab, T* __restrict__ out,
                                        int N, int HW, int C) {
  const long long total = (long long)N * HW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int n = (int)(i / ((long long)HW * C));
    const float* q = ab + ((size_t)n * C + c) * 3;
    out[i] = sg_from_f<T>(fmaf(sg_to_f(z[i]) - q[0], q[1], q[2]));
  }
}

// per-(n,c): style gradients and the coefficients of  g_z = c1*g + k*(z - mean) + c0
__global__ void epi_bwd_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ mr,
                                        const float* __restrict__ style, float* __restrict__ gstyle,
                                        float* __restrict__ coef, int N, int C, int HW) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  int n = i / C, c = i % C;
  float A = (float)sums[(size_t)i * 3], B = (float)sums[(size_t)i * 3 + 1];
  float mean = mr[(size_t)i * 2], rstd = mr[(size_t)i * 2 + 1];
  float s0 = style ? style[(size_t)n * 2 * C + c] : 0.f;
  if (gstyle) {
    gstyle[(size_t)n * 2 * C + c] = B;       // d/ds0 = sum g * xhat
    gstyle[(size_t)n * 2 * C + C + c] = A;   // d/ds1 = sum g
  }
  float inv = 1.f / (float)HW;
  float c1 = rstd * (1.f + s0);
  coef[(size_t)i * 4] = c1;
  coef[(size_t)i * 4 + 1] = -c1 * rstd * B * inv;   // k
  coef[(size_t)i * 4 + 2] = mean;
  coef[(size_t)i * 4 + 3] = -c1 * A * inv;          // c0
}

// same launch scheme as affine_nc_kernel: the 4*VN coefficients of (n, channel vector) live in registers
template <class T>
__global__ void __launch_bounds__(256) epi_bwd_apply_vec_kernel(const T* __restrict__ g, const T* __restrict__ z,
                                                                const float* __restrict__ coef, T* __restrict__ gy,
                                                                int HW, int C, int pix_per_blk, int act) {
  using V = SgVec<T>;
  const int CV = C / V::N;
  const int n = blockIdx.y;
  const int cv = threadIdx.x % CV, pr = threadIdx.x / CV, PR = blockDim.x / CV;
  const int p0 = blockIdx.x * pix_per_blk;
  int p1 = p0 + pix_per_blk;
  if (p1 > HW) p1 = HW;
  float c1[V::N], ck[V::N], cm[V::N], c0[V::N];
#pragma unroll
  for (int k = 0; k < V::N; ++k) {
    const float* q = coef + ((size_t)n * C + cv * V::N + k) * 4;
    c1[k] = q[0]; ck[k] = q[1]; cm[k] = q[2]; c0[k] = q[3];
  }
  const size_t base = ((size_t)n * HW) * C + cv * V::N;
  auto apply = [&](const V& vg, const V& vz, V& o) {
#pragma unroll
    for (int k = 0; k < V::N; ++k) {
      const float zv = vz.get(k);
      const float gz = fmaf(c1[k], vg.get(k), fmaf(ck[k], zv - cm[k], c0[k]));
      o.set(k, (!act || zv > 0.f) ? gz : SG_LRELU_SLOPE * gz);
    }
  };
  constexpr int U = 2;
  int p = p0 + pr;
  for (; p + (U - 1) * PR < p1; p += U * PR) {
    V vg[U], vz[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      vg[u].load(g + base + (size_t)(p + u * PR) * C);
      vz[u].load(z + base + (size_t)(p + u * PR) * C);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      V o;
      apply(vg[u], vz[u], o);
      o.store(gy + base + (size_t)(p + u * PR) * C);
    }
  }
  for (; p < p1; p += PR) {
    V vg, vz, o;
    vg.load(g + base + (size_t)p * C);
    vz.load(z + base + (size_t)p * C);
    apply(vg, vz, o);
    o.store(gy + base + (size_t)p * C);
  }
}

template <class T>
__global__ void epi_bwd_apply_kernel(const T* __restrict__ g, const T* __restrict__ z, const float* __restrict__ coef,
                                     T* __restrict__ gy, int N, int HW, int C, int act) {
  const long long total = (long long)N * HW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int n = (int)(i / ((long long)HW * C));
    const float* q = coef + ((size_t)n * C + c) * 4;
    float zv = sg_to_f(z[i]);
    float gz = fmaf(q[0], sg_to_f(g[i]), fmaf(q[1], zv - q[2], q[3]));
    gy[i] = sg_from_f<T>((!act || zv > 0.f) ? gz : SG_LRELU_SLOPE * gz);
  }
}

// blocks per image of the two-value reductions: >= ~4 waves of CTAs overall, each CTA at least 64 pixels
static void reduce_chunks(int N, int HW, int* chunks_out, int* pix_per_blk_out) {
  int chunks = (148 * 4 + N - 1) / N;
  int maxchunks = (HW + 63) / 64;
  if (chunks > maxchunks) chunks = maxchunks;
  if (chunks < 1) chunks = 1;
  int pix_per_blk = (HW + chunks - 1) / chunks;
  *chunks_out = (HW + pix_per_blk - 1) / pix_per_blk;
  *pix_per_blk_out = pix_per_blk;
}

template <class T, int MODE>
int launch_reduce(const void* in0, const void* in1, void* zout, const float* bias, float bias_mul, const float* noise,
                  const float* noise_w, const float* mr, double* sums, float* ws, int N, int HW, int C, int act,
                  cudaStream_t st) {
  if (ws == nullptr) { sg_set_error("epilogue: workspace missing (sg_epi_ws_floats)"); return 1000; }
  int bx = C >= 128 ? 128 : (C >= 64 ? 64 : 32);
  dim3 block(bx, EPI_ROWS);
  int chunks, pix_per_blk;
  reduce_chunks(N, HW, &chunks, &pix_per_blk);
  size_t smem = (size_t)2 * EPI_ROWS * C * sizeof(float);
  if (smem > 48 * 1024) { sg_set_error("epilogue: C too large"); return 1000; }
  dim3 grid(chunks, N);
  const int VN = SgVec<T>::N;
  if (C % VN == 0 && C / VN <= 256 && (size_t)2 * (256 / (C / VN)) * C * sizeof(float) <= 48 * 1024) {
    size_t smv = (size_t)2 * (256 / (C / VN)) * C * sizeof(float);
    epi_reduce_vec_kernel<T, MODE><<<grid, 256, smv, st>>>((const T*)in0, (const T*)in1, (T*)zout, bias, bias_mul,
                                                           noise, noise_w, mr, sums, ws, HW, C, pix_per_blk, act);
  } else {
    epi_reduce_kernel<T, MODE><<<grid, block, smem, st>>>((const T*)in0, (const T*)in1, (T*)zout, bias, bias_mul, noise,
                                                          noise_w, mr, sums, ws, HW, C, pix_per_blk, act);
  }
  int rc = sg_check_launch("epi_reduce");
  if (rc) return rc;
  epi_sum_parts_kernel<<<sg_cdiv((long long)N * C * 2, 128), 128, 0, st>>>(ws, chunks, sums, N, C);
  return sg_check_launch("epi_sum_parts");
}

}  // namespace

// floats of workspace the two reductions need (block partials)
SG_EXPORT long long sg_epi_ws_floats(int N, int HW, int C) {
  int chunks, ppb;
  reduce_chunks(N, HW, &chunks, &ppb);
  return (long long)N * chunks * 2 * C;
}

// sums: N*C*3 doubles (sum, sum of squares, pivot), fully written (no zeroing needed); ws: sg_epi_ws_floats floats.
SG_EXPORT int sg_epi_act_stats(const void* y, const float* bias, float bias_mul, const float* noise,
                               const float* noise_w, int act, void* z, double* sums, float* ws, int dtype, int N,
                               int HW, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_DISPATCH(dtype, { return launch_reduce<T, 0>(y, nullptr, z, bias, bias_mul, noise, noise_w, nullptr, sums, ws, N, HW, C, act, st); });
  return 0;
}

SG_EXPORT int sg_epi_finalize(const double* sums, const float* style, float* mr, float* ab, int N, int C, int HW,
                              float eps, void* stream) {
  epi_finalize_kernel<<<sg_cdiv((long long)N * C, 128), 128, 0, (cudaStream_t)stream>>>(sums, style, mr, ab, N, C, HW, eps);
  return sg_check_launch("sg_epi_finalize");
}

// (pixel chunks, N) grid for the register-coefficient kernels: >= ~8 CTAs per SM overall, >= 64 pixels per CTA
static void per_image_grid(int N, int HW, int* chunks, int* pix_per_blk) {
  int c = (148 * 8 + N - 1) / N;
  const int maxc = (HW + 63) / 64;
  if (c > maxc) c = maxc;
  if (c < 1) c = 1;
  *pix_per_blk = (HW + c - 1) / c;
  *chunks = (HW + *pix_per_blk - 1) / *pix_per_blk;
}

SG_EXPORT int sg_affine_nc(const void* z, const float* ab, void* out, int dtype, int N, int HW, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_DISPATCH(dtype, {
    const int VN = SgVec<T>::N;
    if (C % VN == 0 && C / VN <= 256 && N <= 65535) {
      const int CV = C / VN;
      int chunks, ppb;
      per_image_grid(N, HW, &chunks, &ppb);
      affine_nc_kernel<T><<<dim3(chunks, N), (256 / CV) * CV, 0, st>>>((const T*)z, ab, (T*)out, HW, C, ppb);
    } else {
      long long total = (long long)N * HW * C;
      affine_nc_scalar_kernel<T><<<sg_grid_1d(total, 256, 148 * 32), 256, 0, st>>>((const T*)z, ab, (T*)out, N, HW, C);
    }
  });
  return sg_check_launch("sg_affine_nc");
}

SG_EXPORT int sg_epi_bwd_reduce(const void* g, const void* z, const float* mr, double* sums, float* ws, int dtype, int N,
                                int HW, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SG_DISPATCH(dtype, { return launch_reduce<T, 1>(g, z, nullptr, nullptr, 0.f, nullptr, nullptr, mr, sums, ws, N, HW, C, 0, st); });
  return 0;
}

SG_EXPORT int sg_epi_bwd_finalize(const double* sums, const float* mr, const float* style, float* gstyle, float* coef,
                                  int N, int C, int HW, void* stream) {
  epi_bwd_finalize_kernel<<<sg_cdiv((long long)N * C, 128), 128, 0, (cudaStream_t)stream>>>(sums, mr, style, gstyle, coef, N, C, HW);
  return sg_check_launch("sg_epi_bwd_finalize");
}

SG_EXPORT int sg_epi_bwd_apply(const void* g, const void* z, const float* coef, int act, void* gy, int dtype, int N,
                               int HW, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  long long total = (long long)N * HW * C;
  SG_DISPATCH(dtype, {
    const int VN = SgVec<T>::N;
    if (C % VN == 0 && C / VN <= 256 && N <= 65535) {
      const int CV = C / VN;
      int chunks, ppb;
      per_image_grid(N, HW, &chunks, &ppb);
      epi_bwd_apply_vec_kernel<T><<<dim3(chunks, N), (256 / CV) * CV, 0, st>>>((const T*)g, (const T*)z, coef, (T*)gy, HW, C, ppb, act);
    } else {
      epi_bwd_apply_kernel<T><<<sg_grid_1d(total, 256, 148 * 32), 256, 0, st>>>((const T*)g, (const T*)z, coef, (T*)gy, N, HW, C, act);
    }
  });
  return sg_check_launch("sg_epi_bwd_apply");
}
