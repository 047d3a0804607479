// Small fp32 kernels: equalized-LR dense layers (mapping MLP, StyleMod affine, discriminator head), PixelNorm,
// MinibatchStdDev (fwd / bwd / bwd-of-bwd for R1), logistic-loss reductions, squared-sum (R1 penalty).
#include "sg_common.cuh"
#include "sg_fast.cuh"
#include "../../include/sgb200.h"

namespace {

// ---------------------------------------------------------------------------------------------------------
// y[b][o] = act(scale * sum_i x[b*ldx + i] * W[o][i] + bias[o]*bias_mul)        reference CustomLayers.py:99-103
// one warp per output feature o.
// ---------------------------------------------------------------------------------------------------------
constexpr int LF_OB = 8;      // outputs per block (one warp each)
constexpr int LF_IC = 256;    // input chunk staged in shared memory
constexpr int LF_BMAX = 32;   // batch rows per pass
// block = 8 warps; warp w owns output o = blockIdx.x*8 + w; x is staged [rows][256] in shared memory so that it is
// read from L2 once per block instead of once per warp.
// gridDim.y > 1 splits the input dimension: slice k writes its partial products to ws[k][B][O] and linear_finish_kernel
// sums the slices in index order (deterministic) and applies scale / bias / activation.
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ x, long long ldx,
                                                         const float* __restrict__ W, const float* __restrict__ bias,
                                                         float scale, float bias_mul, int act, float* __restrict__ y,
                                                         int B, int I, int O, float* __restrict__ ws) {
  __shared__ float xs[LF_BMAX][LF_IC + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * LF_OB + warp;
  const float* w = W + (size_t)(o < O ? o : 0) * I;
  const bool split = gridDim.y > 1;
  const int i_per = ((I + gridDim.y - 1) / gridDim.y + LF_IC - 1) / LF_IC * LF_IC;
  const int i_begin = blockIdx.y * i_per;
  const int i_end = (i_begin + i_per < I) ? (i_begin + i_per) : I;
  for (int b0 = 0; b0 < B; b0 += LF_BMAX) {
    const int nb = (B - b0 < LF_BMAX) ? (B - b0) : LF_BMAX;
    float acc[LF_BMAX];
#pragma unroll
    for (int j = 0; j < LF_BMAX; ++j) acc[j] = 0.f;
    for (int i0 = i_begin; i0 < i_end; i0 += LF_IC) {
      __syncthreads();
      for (int t = threadIdx.x; t < nb * LF_IC; t += blockDim.x) {
        int r = t / LF_IC, c = t % LF_IC;
        xs[r][c] = (i0 + c < I) ? x[(size_t)(b0 + r) * ldx + i0 + c] : 0.f;
      }
      __syncthreads();
      if (o < O) {
#pragma unroll
        for (int cc = 0; cc < LF_IC / 32; ++cc) {
          int c = cc * 32 + lane;
          float wv = (i0 + c < I) ? w[i0 + c] : 0.f;
#pragma unroll
          for (int j = 0; j < LF_BMAX; ++j)
            if (j < nb) acc[j] = fmaf(wv, xs[j][c], acc[j]);
        }
      }
    }
    if (o < O) {
#pragma unroll
      for (int j = 0; j < LF_BMAX; ++j) {
        if (j < nb) {
          float s = sg_warp_sum(acc[j]);
          if (lane == 0) {
            if (split) {
              ws[((size_t)blockIdx.y * B + (b0 + j)) * O + o] = s;
            } else {
              s = s * scale + (bias ? bias[o] * bias_mul : 0.f);
              y[(size_t)(b0 + j) * O + o] = act ? sg_lrelu(s) : s;
            }
          }
        }
      }
    }
  }
}

__global__ void linear_finish_kernel(float* __restrict__ y, const float* __restrict__ ws, int splits,
                                     const float* __restrict__ bias, float scale, float bias_mul, int act, int B, int O) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * O) return;
  float acc = 0.f;
  for (int k = 0; k < splits; ++k) acc += ws[(size_t)k * B * O + i];
  float s = acc * scale + (bias ? bias[i % O] * bias_mul : 0.f);
  y[i] = act ? sg_lrelu(s) : s;
}

// gx[b][i] = scale * sum_o gy[b][o] * W[o][i].  grid (ceil(I/64), ceil(B/32), OS) in clusters (1, 1, OS): a thread owns
// column i for 8 batch rows; CTA z of a cluster walks its share of the 64-wide o chunks (gy staged in shared memory,
// eight independent weight loads in flight), then the OS partial tiles are added in rank order through distributed
// shared memory -- rank r finishes rows r, r + OS, ... -- so every output element has one writer and one summation
// order: no atomics, no workspace, bitwise reproducible.  (Walking all of o in one CTA is a chain of O/8 dependent
// load rounds: 139 us for the 512 x 512 mapping layers; the cluster split cuts the chain by OS.)
constexpr int LD_BB = 32;
constexpr int LD_MAX_SPLIT = 8;
__global__ void __launch_bounds__(256) linear_dgrad_kernel(const float* __restrict__ gy, const float* __restrict__ W,
                                                           float scale, float* __restrict__ gx, int B, int I, int O) {
  __shared__ float gs[LD_BB][65];
  __shared__ float red[LD_BB][64];
  const int OS = gridDim.z, z = blockIdx.z;
  const int il = threadIdx.x & 63, i = blockIdx.x * 64 + il, bq = threadIdx.x >> 6;
  const int b0 = blockIdx.y * LD_BB;
  const int ic = i < I ? i : I - 1;
  const int chunks = (O + 63) / 64, per = (chunks + OS - 1) / OS;
  const int o_begin = min(O, z * per * 64), o_end = min(O, (z + 1) * per * 64);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int oc = o_begin; oc < o_end; oc += 64) {
    __syncthreads();
    for (int e = threadIdx.x; e < LD_BB * 64; e += blockDim.x) {
      const int b = e >> 6, o = e & 63;
      gs[b][o] = (b0 + b < B && oc + o < o_end) ? gy[(size_t)(b0 + b) * O + oc + o] : 0.f;
    }
    __syncthreads();
    const int on = min(64, o_end - oc);
    float wv[8];
    for (int o = 0; o < on; o += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) wv[u] = (o + u < on) ? W[(size_t)(oc + o + u) * I + ic] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(gs[bq * 8 + j][o + u], wv[u], acc[j]);
    }
  }
  if (OS == 1) {
    if (i < I) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int b = b0 + bq * 8 + j;
        if (b < B) gx[(size_t)b * I + i] = acc[j] * scale;
      }
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[bq * 8 + j][il] = acc[j];
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const uint32_t red_base = (uint32_t)__cvta_generic_to_shared(&red[0][0]);
  for (int r = z + OS * bq; r < LD_BB; r += OS * 4) {          // rows r == z (mod OS); the 4 thread quarters interleave
    float s = 0.f;
    for (int k = 0; k < OS; ++k) {
      uint32_t remote;
      float v;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(red_base + (uint32_t)(r * 64 + il) * 4u), "r"(k));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote));
      s += v;
    }
    if (i < I && b0 + r < B) gx[(size_t)(b0 + r) * I + i] = s * scale;
  }
  // nobody may exit while a peer still reads its tile
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// gW[o][i] (+)= scale * sum_b gy[b][o] * x[b*ldx + i];   gb[o] (+)= bias_mul * sum_b gy[b][o]
__global__ void linear_wgrad_kernel(const float* __restrict__ gy, const float* __restrict__ x, long long ldx,
                                    float scale, float* __restrict__ gW, int B, int I, int O, int accumulate) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)O * I) return;
  const int i = (int)(idx % I), o = (int)(idx / I);
  float s = 0.f;
  for (int b = 0; b < B; ++b) s = fmaf(gy[(size_t)b * O + o], x[(size_t)b * ldx + i], s);
  gW[idx] = (accumulate ? gW[idx] : 0.f) + s * scale;
}
__global__ void colsum_kernel(const float* __restrict__ gy, float* __restrict__ gb, int B, int O, float scale,
                              int accumulate) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= O) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += gy[(size_t)b * O + o];
  gb[o] = (accumulate ? gb[o] : 0.f) + s * scale;
}

// ---------------------------------------------------------------------------------------------------------
// PixelNorm over features: y = x * rsqrt(mean(x^2) + eps)                       reference CustomLayers.py:22-23
// one warp per row.
// ---------------------------------------------------------------------------------------------------------
__global__ void pixelnorm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) { float v = x[(size_t)b * D + i]; s = fmaf(v, v, s); }
  s = sg_warp_sum(s);
  float r = rsqrtf(s / D + eps);
  for (int i = lane; i < D; i += 32) y[(size_t)b * D + i] = x[(size_t)b * D + i] * r;
}
// gx = r*g - x * r^3 * mean(g*x)
__global__ void pixelnorm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, float* __restrict__ gx,
                                     int B, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  float s = 0.f, t = 0.f;
  for (int i = lane; i < D; i += 32) {
    float v = x[(size_t)b * D + i];
    s = fmaf(v, v, s);
    t = fmaf(v, g[(size_t)b * D + i], t);
  }
  s = sg_warp_sum(s);
  t = sg_warp_sum(t);
  float r = rsqrtf(s / D + eps);
  float k = r * r * r * t / D;
  for (int i = lane; i < D; i += 32) gx[(size_t)b * D + i] = r * g[(size_t)b * D + i] - k * x[(size_t)b * D + i];
}

// ---------------------------------------------------------------------------------------------------------
// MinibatchStdDev (reference CustomLayers.py:294-305, num_new_features=1) on NHWC fp32 [B, HW, C]:
//   sample b belongs to group m = b % M (M = B/G); y_m = mean_f sqrt(var_j x[j*M+m, f] + 1e-8)
//   out [B, HW, C+1] = concat(x, y_m broadcast).
// ---------------------------------------------------------------------------------------------------------
__global__ void mbstd_stat_kernel(const float* __restrict__ x, float* __restrict__ ym, int G, int M, int F) {
  // one block per group m
  __shared__ float sm[32];
  const int m = blockIdx.x;
  float acc = 0.f;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float mu = 0.f;
    for (int j = 0; j < G; ++j) mu += x[(size_t)(j * M + m) * F + f];
    mu /= G;
    float v = 0.f;
    for (int j = 0; j < G; ++j) { float c = x[(size_t)(j * M + m) * F + f] - mu; v = fmaf(c, c, v); }
    acc += sqrtf(v / G + 1e-8f);
  }
  acc = sg_block_sum(acc, sm);
  if (threadIdx.x == 0) ym[m] = acc / F;
}
__global__ void mbstd_concat_kernel(const float* __restrict__ x, const float* __restrict__ ym, float* __restrict__ out,
                                    int B, int HW, int C, int M) {
  const long long total = (long long)B * HW * (C + 1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % (C + 1));
    long long p = i / (C + 1);
    int b = (int)(p / HW);
    out[i] = c < C ? x[(size_t)p * C + c] : ym[b % M];
  }
}
// Gm[m] = sum over the group's samples and pixels of g_out[..., C]
__global__ void mbstd_gsum_kernel(const float* __restrict__ gout, float* __restrict__ Gm, int G, int M, int HW, int C) {
  __shared__ float sm[32];
  const int m = blockIdx.x;
  float acc = 0.f;
  for (int t = threadIdx.x; t < G * HW; t += blockDim.x) {
    int j = t / HW, p = t % HW;
    acc += gout[((size_t)(j * M + m) * HW + p) * (C + 1) + C];
  }
  acc = sg_block_sum(acc, sm);
  if (threadIdx.x == 0) Gm[m] = acc;
}
// gx[b,f] = gout_x[b,f] + Gm[m] * c_bf / (F*G*sigma_f)
__global__ void mbstd_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ x,
                                 const float* __restrict__ Gm, float* __restrict__ gx, int G, int M, int HW, int C) {
  const int F = HW * C;
  const long long total = (long long)M * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int f = (int)(i % F);
    int m = (int)(i / F);
    int p = f / C, c = f % C;
    float mu = 0.f;
    for (int j = 0; j < G; ++j) mu += x[(size_t)(j * M + m) * F + f];
    mu /= G;
    float v = 0.f;
    for (int j = 0; j < G; ++j) { float d = x[(size_t)(j * M + m) * F + f] - mu; v = fmaf(d, d, v); }
    float sig = sqrtf(v / G + 1e-8f);
    float k = Gm[m] / ((float)F * G * sig);
    for (int j = 0; j < G; ++j) {
      size_t b = (size_t)(j * M + m);
      float go = gout ? gout[(b * HW + p) * (C + 1) + c] : 0.f;
      gx[b * F + f] = go + k * (x[b * F + f] - mu);
    }
  }
}
// Double backward of the map (x, gout) -> gx.  Given cotangent t[b,f] for gx:
//   cot_gout[b,p,c<C] = t;  cot_gout[b,p,C] = Tm[m] = sum_{j,f} t*c/(F*G*sigma)
//   cot_x[k,f] = Gm/(F*G*sigma) * ( t_k - mean_j t_j - c_k * mean_j(t_j c_j) / sigma^2 )
__global__ void mbstd_bwd2_tm_kernel(const float* __restrict__ t, const float* __restrict__ x, float* __restrict__ Tm,
                                     int G, int M, int F) {
  __shared__ float sm[32];
  const int m = blockIdx.x;
  float acc = 0.f;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float mu = 0.f;
    for (int j = 0; j < G; ++j) mu += x[(size_t)(j * M + m) * F + f];
    mu /= G;
    float v = 0.f, tc = 0.f;
    for (int j = 0; j < G; ++j) {
      float d = x[(size_t)(j * M + m) * F + f] - mu;
      v = fmaf(d, d, v);
      tc = fmaf(t[(size_t)(j * M + m) * F + f], d, tc);
    }
    acc += tc / sqrtf(v / G + 1e-8f);
  }
  acc = sg_block_sum(acc, sm);
  if (threadIdx.x == 0) Tm[m] = acc / ((float)F * G);
}
__global__ void mbstd_bwd2_kernel(const float* __restrict__ t, const float* __restrict__ x,
                                  const float* __restrict__ Gm, const float* __restrict__ Tm,
                                  float* __restrict__ cot_gout, float* __restrict__ cot_x, int G, int M, int HW, int C) {
  const int F = HW * C;
  const long long total = (long long)M * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int f = (int)(i % F);
    int m = (int)(i / F);
    int p = f / C, c = f % C;
    float mu = 0.f, tm = 0.f;
    for (int j = 0; j < G; ++j) { mu += x[(size_t)(j * M + m) * F + f]; tm += t[(size_t)(j * M + m) * F + f]; }
    mu /= G; tm /= G;
    float v = 0.f, tc = 0.f;
    for (int j = 0; j < G; ++j) {
      float d = x[(size_t)(j * M + m) * F + f] - mu;
      v = fmaf(d, d, v);
      tc = fmaf(t[(size_t)(j * M + m) * F + f], d, tc);
    }
    tc /= G;
    float s2 = v / G + 1e-8f;
    float sig = sqrtf(s2);
    float k = Gm[m] / ((float)F * G * sig);
    for (int j = 0; j < G; ++j) {
      size_t b = (size_t)(j * M + m);
      float tv = t[b * F + f];
      cot_x[b * F + f] = k * (tv - tm - (x[b * F + f] - mu) * tc / s2);
      cot_gout[(b * HW + p) * (C + 1) + c] = tv;
      if (c == 0) cot_gout[(b * HW + p) * (C + 1) + C] = Tm[m];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Logistic-loss pieces (reference Losses.py:213-229): out[0] (+)= scale * sum_i softplus(sign * x[i]);
// backward: gx[i] = gscale * sign * sigmoid(sign * x[i]).   Single block (B is tiny).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float softplus_f(float v) { return v > 20.f ? v : log1pf(expf(v)); }
__global__ void softplus_sum_kernel(const float* __restrict__ x, int n, float sign, float scale,
                                    float* __restrict__ out, int accumulate) {
  __shared__ float sm[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += softplus_f(sign * x[i]);
  s = sg_block_sum(s, sm);
  if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + s * scale;
}
__global__ void softplus_bwd_kernel(const float* __restrict__ x, int n, float sign, const float* __restrict__ gout,
                                    float gscale, float* __restrict__ gx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = sign * x[i];
  float sg = 1.f / (1.f + expf(-v));
  gx[i] = gout[0] * gscale * sign * sg;
}

// out[0] (+)= scale * sum x^2 (fp32 data, fp64 accumulation across blocks via partials)
__global__ void sqsum_partial_kernel(const float* __restrict__ x, long long n, double* __restrict__ part) {
  __shared__ float sm[32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    s = fmaf(v, v, s);
  }
  s = sg_block_sum(s, sm);
  if (threadIdx.x == 0) part[blockIdx.x] = (double)s;
}
__global__ void sqsum_final_kernel(const double* __restrict__ part, int nblk, float scale, float* __restrict__ out,
                                   int accumulate) {
  __shared__ double sd[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) s += part[i];
  sd[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sd[threadIdx.x] += sd[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + (float)(sd[0] * scale);
}
// gx = (gout[0]*coef) * x
__global__ void scale_by_scalar_kernel(const float* __restrict__ x, const float* __restrict__ gout, float coef,
                                       float* __restrict__ gx, long long n) {
  float k = gout[0] * coef;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    gx[i] = k * x[i];
}

}  // namespace

static int linear_fwd_splits(int I) {
  int splits = 1;
  if (I >= 2048) {                                  // wide input (discriminator dense0): split I to fill the GPU
    splits = I / 1024;
    if (splits > 16) splits = 16;
  }
  return splits;
}
// floats of workspace sg_linear_fwd needs for this shape (0: none)
SG_EXPORT long long sg_linear_fwd_ws_floats(int B, int I, int O) {
  const int splits = linear_fwd_splits(I);
  return splits > 1 ? (long long)splits * B * O : 0;
}
SG_EXPORT int sg_linear_fwd(const float* x, long long ldx, const float* W, const float* bias, float scale,
                            float bias_mul, int act, float* y, int B, int I, int O, float* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (g_sg_fast_paths) {               // I in {128, 256, 512}: weights-in-registers kernel of csrc/sg_linear_group.cu
    int rc = sgf_linear_fwd(x, ldx, W, bias, scale, bias_mul, act, y, B, I, O, stream);
    if (rc != SG_FAST_UNSUPPORTED) { if (rc) sg_set_error("sg_linear_fwd/group: launch failed"); return rc; }
  }
  const int splits = linear_fwd_splits(I);
  SG_REQUIRE(splits == 1 || ws != nullptr, "sg_linear_fwd: workspace missing (sg_linear_fwd_ws_floats)");
  linear_fwd_kernel<<<dim3(sg_cdiv(O, LF_OB), splits), 256, 0, st>>>(x, ldx, W, bias, scale, bias_mul, act, y, B, I, O, ws);
  if (splits > 1)
    linear_finish_kernel<<<sg_cdiv((long long)B * O, 256), 256, 0, st>>>(y, ws, splits, bias, scale, bias_mul, act, B, O);
  return sg_check_launch("sg_linear_fwd");
}

SG_EXPORT int sg_linear_dgrad(const float* gy, const float* W, float scale, float* gx, int B, int I, int O,
                              void* stream) {
  int os = sg_cdiv(O, 64);
  if (os > LD_MAX_SPLIT) os = LD_MAX_SPLIT;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sg_cdiv(I, 64), sg_cdiv(B, LD_BB), os);
  cfg.blockDim = dim3(256);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = os;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, linear_dgrad_kernel, gy, W, scale, gx, B, I, O);
  if (e != cudaSuccess) { sg_set_error("sg_linear_dgrad: launch failed"); return (int)e; }
  return sg_check_launch("sg_linear_dgrad");
}

SG_EXPORT int sg_linear_wgrad(const float* gy, const float* x, long long ldx, float scale, float bias_scale,
                              float* gW, float* gb, int B, int I, int O, int accumulate, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (g_sg_fast_paths && gW) {
    int rc = sgf_linear_wgrad(gy, x, ldx, scale, bias_scale, gW, gb, B, I, O, accumulate, stream);
    if (rc != SG_FAST_UNSUPPORTED) { if (rc) sg_set_error("sg_linear_wgrad/group: launch failed"); return rc; }
  }
  if (gW) {
    linear_wgrad_kernel<<<sg_cdiv((long long)O * I, 256), 256, 0, st>>>(gy, x, ldx, scale, gW, B, I, O, accumulate);
    int rc = sg_check_launch("sg_linear_wgrad");
    if (rc) return rc;
  }
  if (gb) {
    colsum_kernel<<<sg_cdiv(O, 128), 128, 0, st>>>(gy, gb, B, O, bias_scale, accumulate);
    return sg_check_launch("sg_linear_wgrad/bias");
  }
  return 0;
}

SG_EXPORT int sg_pixelnorm_fwd(const float* x, float* y, int B, int D, float eps, void* stream) {
  pixelnorm_fwd_kernel<<<sg_cdiv(B, 4), 128, 0, (cudaStream_t)stream>>>(x, y, B, D, eps);
  return sg_check_launch("sg_pixelnorm_fwd");
}
SG_EXPORT int sg_pixelnorm_bwd(const float* g, const float* x, float* gx, int B, int D, float eps, void* stream) {
  pixelnorm_bwd_kernel<<<sg_cdiv(B, 4), 128, 0, (cudaStream_t)stream>>>(g, x, gx, B, D, eps);
  return sg_check_launch("sg_pixelnorm_bwd");
}

// ws: at least 2*M floats
SG_EXPORT int sg_mbstd_fwd(const float* x, float* out, float* ws, int B, int HW, int C, int G, void* stream) {
  SG_REQUIRE(G >= 1 && B % G == 0, "sg_mbstd_fwd: batch must be divisible by the group size");
  cudaStream_t st = (cudaStream_t)stream;
  int M = B / G, F = HW * C;
  mbstd_stat_kernel<<<M, 256, 0, st>>>(x, ws, G, M, F);
  mbstd_concat_kernel<<<sg_grid_1d((long long)B * HW * (C + 1), 256), 256, 0, st>>>(x, ws, out, B, HW, C, M);
  return sg_check_launch("sg_mbstd_fwd");
}
SG_EXPORT int sg_mbstd_bwd(const float* gout, const float* x, float* gx, float* ws, int B, int HW, int C, int G,
                           void* stream) {
  SG_REQUIRE(G >= 1 && B % G == 0, "sg_mbstd_bwd: batch must be divisible by the group size");
  cudaStream_t st = (cudaStream_t)stream;
  int M = B / G;
  mbstd_gsum_kernel<<<M, 128, 0, st>>>(gout, ws, G, M, HW, C);
  mbstd_bwd_kernel<<<sg_grid_1d((long long)M * HW * C, 256), 256, 0, st>>>(gout, x, ws, gx, G, M, HW, C);
  return sg_check_launch("sg_mbstd_bwd");
}
// t: cotangent of gx [B,HW,C]; gout: the original upstream grad [B,HW,C+1]; outputs cot_gout [B,HW,C+1], cot_x.
SG_EXPORT int sg_mbstd_bwd2(const float* t, const float* gout, const float* x, float* cot_gout, float* cot_x,
                            float* ws, int B, int HW, int C, int G, void* stream) {
  SG_REQUIRE(G >= 1 && B % G == 0, "sg_mbstd_bwd2: batch must be divisible by the group size");
  cudaStream_t st = (cudaStream_t)stream;
  int M = B / G, F = HW * C;
  mbstd_gsum_kernel<<<M, 128, 0, st>>>(gout, ws, G, M, HW, C);
  mbstd_bwd2_tm_kernel<<<M, 256, 0, st>>>(t, x, ws + M, G, M, F);
  mbstd_bwd2_kernel<<<sg_grid_1d((long long)M * F, 256), 256, 0, st>>>(t, x, ws, ws + M, cot_gout, cot_x, G, M, HW, C);
  return sg_check_launch("sg_mbstd_bwd2");
}

SG_EXPORT int sg_softplus_sum(const float* x, int n, float sign, float scale, float* out, int accumulate,
                              void* stream) {
  softplus_sum_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(x, n, sign, scale, out, accumulate);
  return sg_check_launch("sg_softplus_sum");
}
SG_EXPORT int sg_softplus_bwd(const float* x, int n, float sign, const float* gout, float gscale, float* gx,
                              void* stream) {
  softplus_bwd_kernel<<<sg_cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(x, n, sign, gout, gscale, gx);
  return sg_check_launch("sg_softplus_bwd");
}

// ws: 1024 doubles
SG_EXPORT int sg_sqsum(const float* x, long long n, float scale, float* out, int accumulate, double* ws,
                       void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = sg_grid_1d(n, 256, 1024);
  sqsum_partial_kernel<<<nblk, 256, 0, st>>>(x, n, ws);
  sqsum_final_kernel<<<1, 256, 0, st>>>(ws, nblk, scale, out, accumulate);
  return sg_check_launch("sg_sqsum");
}
SG_EXPORT int sg_scale_by_scalar(const float* x, const float* gout, float coef, float* gx, long long n, void* stream) {
  scale_by_scalar_kernel<<<sg_grid_1d(n, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(x, gout, coef, gx, n);
  return sg_check_launch("sg_scale_by_scalar");
}
