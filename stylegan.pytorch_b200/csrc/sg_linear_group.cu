// Grouped dense layers: G independent affine maps  y_g[B,O_g] = scale_g * x_g[B,I] W_g[O_g,I]^T + bias_g * bmul_g  in ONE
// launch, with their data gradient and weight/bias gradients as one launch each.
//
// reference call sites: the 2 x (number of synthesis layers) StyleMod affines of one Generator forward
// (models/CustomLayers.py:203-216: EqualizedLinear(latent, 2C) applied to dlatents_in_range[:, l]) -- 14 at 256x256, 18 at
// 1024x1024 -- which the reference (and round 1 of this engine) evaluates as 14 separate tiny GEMMs (38 us each on a
// B200, latency-bound: 16-128 blocks, one dependent global-load chain per warp).  All of them depend only on the dlatents,
// so they run before the synthesis network starts; their gradients arrive together at the end of its backward.
// fp32 throughout (these feed the fp32-parity path as well): CUDA-core FMAs, all global loads of a warp issued up front.
#include "sg_common.cuh"
#include "../../include/sgb200.h"
#include <mutex>
#include <string.h>

namespace {

constexpr int LG_B = 32;        // batch rows per launch (callers with more rows loop)
constexpr int LG_OPB = 16;      // outputs per block (2 per warp, 8 warps)

// One table row per group (device memory, 8 x int64): pointers as integers, floats as their bit patterns.
struct LinGroup {
  long long x;        // const float*  [B, I] with row stride ldx
  long long ldx;
  long long W;        // const float*  [O, I]
  long long bias;     // const float*  [O] or 0
  long long y;        // float*        [B, O]   (fwd: output; bwd: the incoming gradient gy)
  long long O;
  long long scales;   // low 32 bits: scale, high 32 bits: bias_mul
  long long blk0;     // first block of this group in the launch (prefix sum of ceil(O / outputs-per-block))
};

__device__ __forceinline__ float lo_f(long long v) { return __int_as_float((int)(v & 0xffffffffll)); }
__device__ __forceinline__ float hi_f(long long v) { return __int_as_float((int)((unsigned long long)v >> 32)); }

__device__ __forceinline__ int find_group(const LinGroup* __restrict__ tab, int G, int blk) {
  int g = 0;
  while (g + 1 < G && (int)tab[g + 1].blk0 <= blk) ++g;
  return g;
}

// ---- forward: block = 16 outputs of one group; x_g staged once in shared memory -------------------------------------------
template <int IV>   // I = 128 * IV floats (IV float4 per lane)
__global__ void __launch_bounds__(256) lingroup_fwd_kernel(const LinGroup* __restrict__ tab, const LinGroup one, int G,
                                                           int B, int act) {
  extern __shared__ float xs[];                        // [B][I]
  constexpr int I = 128 * IV;
  const LinGroup t = tab ? tab[find_group(tab, G, blockIdx.x)] : one;   // one group passed by value: no table upload
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int O = (int)t.O;
  const int o0 = ((int)blockIdx.x - (int)t.blk0) * LG_OPB + warp * 2;
  // this warp's two weight rows: every load of the warp is in flight before anything is consumed
  const float* Wp = reinterpret_cast<const float*>(t.W);
  float4 w[2][IV];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int o = o0 + r < O ? o0 + r : O - 1;
#pragma unroll
    for (int v = 0; v < IV; ++v) w[r][v] = *reinterpret_cast<const float4*>(Wp + (size_t)o * I + v * 128 + lane * 4);
  }
  const float* xp = reinterpret_cast<const float*>(t.x);
  const int B8 = (B + 7) & ~7;                         // rows are consumed eight at a time: pad with zeros
  for (int i = threadIdx.x; i < B8 * (I / 4); i += blockDim.x) {
    const int b = i / (I / 4), c = i % (I / 4);
    reinterpret_cast<float4*>(xs)[i] = b < B ? *reinterpret_cast<const float4*>(xp + (size_t)b * t.ldx + c * 4)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const float scale = lo_f(t.scales), bmul = hi_f(t.scales);
  const float* bp = reinterpret_cast<const float*>(t.bias);
  float* yp = reinterpret_cast<float*>(t.y);
  for (int b0 = 0; b0 < B; b0 += 8) {                  // 8 batch rows x 2 outputs = 16 accumulators at a time
    float acc[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.f;
#pragma unroll
    for (int v = 0; v < IV; ++v) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + (size_t)(b0 + j) * I + v * 128 + lane * 4);
        acc[j][0] = fmaf(xv.x, w[0][v].x, fmaf(xv.y, w[0][v].y, fmaf(xv.z, w[0][v].z, fmaf(xv.w, w[0][v].w, acc[j][0]))));
        acc[j][1] = fmaf(xv.x, w[1][v].x, fmaf(xv.y, w[1][v].y, fmaf(xv.z, w[1][v].z, fmaf(xv.w, w[1][v].w, acc[j][1]))));
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float s = sg_warp_sum(acc[j][r]);
        if (lane == 0 && o0 + r < O && b0 + j < B) {
          float v = s * scale + (bp ? bp[o0 + r] * bmul : 0.f);
          yp[(size_t)(b0 + j) * O + o0 + r] = act ? sg_lrelu(v) : v;
        }
      }
    }
  }
}

// ---- data gradient: gx_g[b, i] = scale_g * sum_o gy_g[b, o] W_g[o, i];  block = (group, 64 input columns) ----------------
// gx is written with row stride ldx into the tensor x points at (for the StyleMod group: the gradient of dlatents[:, l]).
// grid (G * I/64, OS) in clusters (1, OS): CTA y of a cluster walks its share of the group's 64-wide o chunks, then the OS
// partial tiles are added in rank order through distributed shared memory (rank r finishes rows r, r + OS, ...): one
// writer and one summation order per element, and the chain of dependent weight-load rounds is OS times shorter.
constexpr int LGD_SPLIT = 8;
__global__ void __launch_bounds__(256) lingroup_dgrad_kernel(const LinGroup* __restrict__ tab, int G, int B, int I,
                                                             int accumulate) {
  __shared__ float gs[LG_B][65];                       // gy chunk [b][o within chunk of 64]
  __shared__ float red[LG_B][64];
  const int col_blocks = I / 64;
  const int g = blockIdx.x / col_blocks, cb = blockIdx.x % col_blocks;
  const int OS = gridDim.y, z = blockIdx.y;
  const LinGroup t = tab[g];
  const int O = (int)t.O;
  const int il = threadIdx.x & 63, i = cb * 64 + il, bq = threadIdx.x >> 6;      // 4 batch quarters of 8 rows
  const float* Wp = reinterpret_cast<const float*>(t.W);
  const float* gyp = reinterpret_cast<const float*>(t.y);
  const int chunks = (O + 63) / 64, per = (chunks + OS - 1) / OS;
  const int o_begin = min(O, z * per * 64), o_end = min(O, (z + 1) * per * 64);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int oc = o_begin; oc < o_end; oc += 64) {
    __syncthreads();
    for (int e = threadIdx.x; e < LG_B * 64; e += blockDim.x) {
      const int b = e >> 6, o = e & 63;
      gs[b][o] = (b < B && oc + o < o_end) ? gyp[(size_t)b * O + oc + o] : 0.f;
    }
    __syncthreads();
    const int on = min(64, o_end - oc);
    float wv[8];
    for (int o = 0; o < on; o += 8) {                  // 8 independent weight loads per thread in flight
#pragma unroll
      for (int u = 0; u < 8; ++u) wv[u] = (o + u < on) ? Wp[(size_t)(oc + o + u) * I + i] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(gs[bq * 8 + j][o + u], wv[u], acc[j]);
    }
  }
  const float scale = lo_f(t.scales);
  float* gx = reinterpret_cast<float*>(t.x);
#pragma unroll
  for (int j = 0; j < 8; ++j) red[bq * 8 + j][il] = acc[j];
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const uint32_t red_base = (uint32_t)__cvta_generic_to_shared(&red[0][0]);
  for (int r = z + OS * bq; r < LG_B; r += OS * 4) {
    float s = 0.f;
    for (int k = 0; k < OS; ++k) {
      uint32_t remote;
      float v;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(red_base + (uint32_t)(r * 64 + il) * 4u), "r"(k));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote));
      s += v;
    }
    if (r < B) {
      float* d = gx + (size_t)r * t.ldx + i;
      *d = (accumulate ? *d : 0.f) + s * scale;
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- weight / bias gradient: gW_g[o, i] = scale_g * sum_b gy_g[b, o] x_g[b, i]; gb_g[o] = bmul_g * sum_b gy_g[b, o] --------
// table rows here: x = activations, W = gW (output), bias = gb (output or 0), y = gy.  block = 8 output rows of one group.
template <int IV>
__global__ void __launch_bounds__(256) lingroup_wgrad_kernel(const LinGroup* __restrict__ tab, const LinGroup one, int G,
                                                             int B, int accumulate) {
  extern __shared__ float xs[];                        // [B][I]
  constexpr int I = 128 * IV;
  __shared__ float gys[LG_B][8];
  const LinGroup t = tab ? tab[find_group(tab, G, blockIdx.x)] : one;
  const int O = (int)t.O;
  const int o0 = ((int)blockIdx.x - (int)t.blk0) * 8;
  const float* xp = reinterpret_cast<const float*>(t.x);
  const float* gyp = reinterpret_cast<const float*>(t.y);
  for (int i = threadIdx.x; i < B * (I / 4); i += blockDim.x) {
    const int b = i / (I / 4), c = i % (I / 4);
    reinterpret_cast<float4*>(xs)[i] = *reinterpret_cast<const float4*>(xp + (size_t)b * t.ldx + c * 4);
  }
  for (int e = threadIdx.x; e < LG_B * 8; e += blockDim.x) {
    const int b = e >> 3, r = e & 7;
    gys[b][r] = (b < B && o0 + r < O) ? gyp[(size_t)b * O + o0 + r] : 0.f;
  }
  __syncthreads();
  const float scale = lo_f(t.scales), bmul = hi_f(t.scales);
  float* gW = reinterpret_cast<float*>(t.W);
  // thread owns I/256 float2... simpler: each thread handles columns c = tid*4 .. +3 for IV == 4 (I = 512: 128 threads x 4)
  for (int c = threadIdx.x * 4; c < I; c += blockDim.x * 4) {
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    for (int b = 0; b < B; ++b) {
      const float4 xv = *reinterpret_cast<const float4*>(xs + (size_t)b * I + c);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float gv = gys[b][r];
        acc[r][0] = fmaf(gv, xv.x, acc[r][0]); acc[r][1] = fmaf(gv, xv.y, acc[r][1]);
        acc[r][2] = fmaf(gv, xv.z, acc[r][2]); acc[r][3] = fmaf(gv, xv.w, acc[r][3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (o0 + r < O) {
        float4* d = reinterpret_cast<float4*>(gW + (size_t)(o0 + r) * I + c);
        float4 v = make_float4(acc[r][0] * scale, acc[r][1] * scale, acc[r][2] * scale, acc[r][3] * scale);
        if (accumulate) { const float4 old = *d; v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
        *d = v;
      }
  }
  float* gb = reinterpret_cast<float*>(t.bias);
  if (gb && threadIdx.x < 8 && o0 + (int)threadIdx.x < O) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += gys[b][threadIdx.x];
    gb[o0 + threadIdx.x] = (accumulate ? gb[o0 + threadIdx.x] : 0.f) + s * bmul;
  }
}

}  // namespace

SG_EXPORT int sg_lingroup_rows(void) { return 8; }   // int64 words per table row

// blocks of the forward / weight-gradient launch for one group of O outputs (the caller builds the blk0 prefix sums)
SG_EXPORT int sg_lingroup_fwd_blocks(int O) { return (O + LG_OPB - 1) / LG_OPB; }
SG_EXPORT int sg_lingroup_wgrad_blocks(int O) { return (O + 7) / 8; }

SG_EXPORT int sg_lingroup_supported(int B, int I) { return (B >= 1 && B <= LG_B && (I == 512 || I == 256 || I == 128)) ? 1 : 0; }

SG_EXPORT int sg_lingroup_fwd(const void* table, int G, int total_blocks, int B, int I, int act, void* stream) {
  SG_REQUIRE(sg_lingroup_supported(B, I), "sg_lingroup_fwd: B must be <= 32 and I in {128, 256, 512}");
  const LinGroup* tab = reinterpret_cast<const LinGroup*>(table);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)((B + 7) & ~7) * I * sizeof(float);
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(lingroup_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(lingroup_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  });
  const LinGroup none = {};
  if (I == 512) lingroup_fwd_kernel<4><<<total_blocks, 256, smem, st>>>(tab, none, G, B, act);
  else if (I == 256) lingroup_fwd_kernel<2><<<total_blocks, 256, smem, st>>>(tab, none, G, B, act);
  else lingroup_fwd_kernel<1><<<total_blocks, 256, smem, st>>>(tab, none, G, B, act);
  return sg_check_launch("sg_lingroup_fwd");
}

SG_EXPORT int sg_lingroup_dgrad(const void* table, int G, int B, int I, int accumulate, void* stream) {
  SG_REQUIRE(sg_lingroup_supported(B, I), "sg_lingroup_dgrad: B must be <= 32 and I in {128, 256, 512}");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G * (I / 64), LGD_SPLIT);
  cfg.blockDim = dim3(256);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = LGD_SPLIT; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, lingroup_dgrad_kernel, reinterpret_cast<const LinGroup*>(table), G, B, I, accumulate);
  if (e != cudaSuccess) { sg_set_error("sg_lingroup_dgrad: launch failed"); return (int)e; }
  return sg_check_launch("sg_lingroup_dgrad");
}

SG_EXPORT int sg_lingroup_wgrad(const void* table, int G, int total_blocks, int B, int I, void* stream) {
  SG_REQUIRE(sg_lingroup_supported(B, I), "sg_lingroup_wgrad: B must be <= 32 and I in {128, 256, 512}");
  const LinGroup* tab = reinterpret_cast<const LinGroup*>(table);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)B * I * sizeof(float);
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(lingroup_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  });
  const LinGroup none = {};
  if (I == 512) lingroup_wgrad_kernel<4><<<total_blocks, 256, smem, st>>>(tab, none, G, B, 0);
  else if (I == 256) lingroup_wgrad_kernel<2><<<total_blocks, 256, smem, st>>>(tab, none, G, B, 0);
  else lingroup_wgrad_kernel<1><<<total_blocks, 256, smem, st>>>(tab, none, G, B, 0);
  return sg_check_launch("sg_lingroup_wgrad");
}

// ---- single dense layer through the same kernels (sg_linear_fwd / sg_linear_wgrad fast paths, csrc/sg_small.cu): the
//      group descriptor travels as a kernel argument, no table in device memory --------------------------------------------
static long long pack2(float a, float b) {
  unsigned int ua, ub;
  memcpy(&ua, &a, 4); memcpy(&ub, &b, 4);
  return (long long)(((unsigned long long)ub << 32) | ua);
}
static void lg_attrs() {
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(lingroup_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(lingroup_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  });
}
// -> 0 launched, 1000 not applicable (caller keeps its general kernel)
int sgf_linear_fwd(const float* x, long long ldx, const float* W, const float* bias, float scale, float bias_mul, int act,
                   float* y, int B, int I, int O, void* stream) {
  if (!(I == 512 || I == 256 || I == 128) || ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(x) & 15)) return 1000;
  lg_attrs();
  cudaStream_t st = (cudaStream_t)stream;
  for (int b0 = 0; b0 < B; b0 += LG_B) {
    const int nb = B - b0 < LG_B ? B - b0 : LG_B;
    LinGroup t;
    t.x = (long long)(x + (size_t)b0 * ldx); t.ldx = ldx; t.W = (long long)W; t.bias = (long long)bias;
    t.y = (long long)(y + (size_t)b0 * O); t.O = O; t.scales = pack2(scale, bias_mul); t.blk0 = 0;
    const int blocks = (O + LG_OPB - 1) / LG_OPB;
    const size_t smem = (size_t)((nb + 7) & ~7) * I * sizeof(float);
    if (I == 512) lingroup_fwd_kernel<4><<<blocks, 256, smem, st>>>(nullptr, t, 1, nb, act);
    else if (I == 256) lingroup_fwd_kernel<2><<<blocks, 256, smem, st>>>(nullptr, t, 1, nb, act);
    else lingroup_fwd_kernel<1><<<blocks, 256, smem, st>>>(nullptr, t, 1, nb, act);
  }
  return (int)cudaGetLastError();
}
int sgf_linear_wgrad(const float* gy, const float* x, long long ldx, float scale, float bias_scale, float* gW, float* gb,
                     int B, int I, int O, int accumulate, void* stream) {
  if (!(I == 512 || I == 256 || I == 128) || ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) || gW == nullptr)
    return 1000;
  lg_attrs();
  cudaStream_t st = (cudaStream_t)stream;
  for (int b0 = 0; b0 < B; b0 += LG_B) {
    const int nb = B - b0 < LG_B ? B - b0 : LG_B;
    LinGroup t;
    t.x = (long long)(x + (size_t)b0 * ldx); t.ldx = ldx; t.W = (long long)gW; t.bias = (long long)gb;
    t.y = (long long)(gy + (size_t)b0 * O); t.O = O; t.scales = pack2(scale, bias_scale); t.blk0 = 0;
    const int blocks = (O + 7) / 8;
    const size_t smem = (size_t)nb * I * sizeof(float);
    const int acc = (accumulate || b0 > 0) ? 1 : 0;
    if (I == 512) lingroup_wgrad_kernel<4><<<blocks, 256, smem, st>>>(nullptr, t, 1, nb, acc);
    else if (I == 256) lingroup_wgrad_kernel<2><<<blocks, 256, smem, st>>>(nullptr, t, 1, nb, acc);
    else lingroup_wgrad_kernel<1><<<blocks, 256, smem, st>>>(nullptr, t, 1, nb, acc);
  }
  return (int)cudaGetLastError();
}
