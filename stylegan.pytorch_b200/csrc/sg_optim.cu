// Multi-tensor optimizer step (Adam + global-norm clip + 1/world gradient scale + generator EMA), gradient
// square-norm, and the Philox normal generator used by NoiseLayer.
// reference: GAN.py:529-533 (Adam), GAN.py:651 (clip_grad_norm_), models/__init__.py:13-40 (update_average),
//            CustomLayers.py:193 (torch.randn noise).
#include "sg_common.cuh"
#include "../../include/sgb200.h"

namespace {

constexpr int CHUNK = 2048;  // elements per block

__device__ __forceinline__ int find_entry(const SgTensorEntry* __restrict__ tab, int T, int chunk) {
  int lo = 0, hi = T - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (tab[mid].chunk0 <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// stage 1: one partial per chunk; stage 2 (one block) sums them in index order -> bitwise reproducible, no atomics
__global__ void sqnorm_multi_kernel(const SgTensorEntry* __restrict__ tab, int T, double* __restrict__ partial) {
  __shared__ float sm[32];
  const int e = find_entry(tab, T, blockIdx.x);
  const SgTensorEntry en = tab[e];
  const float* g = reinterpret_cast<const float*>(en.g);
  long long i0 = (long long)(blockIdx.x - en.chunk0) * CHUNK;
  long long i1 = i0 + CHUNK < en.n ? i0 + CHUNK : en.n;
  float s = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) { float v = g[i]; s = fmaf(v, v, s); }
  s = sg_block_sum(s, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = (double)s;
}
__global__ void sqnorm_final_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  __shared__ double sd[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  sd[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sd[threadIdx.x] += sd[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sd[0];
}

__global__ void adam_multi_kernel(const SgTensorEntry* __restrict__ tab, int T, float lr, float beta1, float beta2,
                                  float eps, float grad_scale, const double* __restrict__ normsq, float max_norm,
                                  float ema_beta) {
  const int e = find_entry(tab, T, blockIdx.x);
  const SgTensorEntry en = tab[e];
  float* p = reinterpret_cast<float*>(en.p);
  const float* g = reinterpret_cast<const float*>(en.g);
  float* m = reinterpret_cast<float*>(en.m);
  float* v = reinterpret_cast<float*>(en.v);
  float* ema = reinterpret_cast<float*>(en.ema);
  const float t = *reinterpret_cast<const float*>(en.step) + 1.f;
  const float bc1 = 1.f - powf(beta1, t);
  const float bc2 = 1.f - powf(beta2, t);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  float gs = grad_scale;
  if (normsq != nullptr) {
    // clip_grad_norm_: coef = min(1, max_norm / (||g|| + 1e-6)) on the already-scaled gradient
    float total = (float)sqrt(*normsq) * grad_scale;
    float coef = max_norm / (total + 1e-6f);
    if (coef < 1.f) gs *= coef;
  }
  long long i0 = (long long)(blockIdx.x - en.chunk0) * CHUNK;
  long long i1 = i0 + CHUNK < en.n ? i0 + CHUNK : en.n;
  auto update = [&](float gvr, float& mr, float& vr, float& pr, float& er) {
    const float gv = gvr * gs;
    mr = beta1 * mr + (1.f - beta1) * gv;
    vr = beta2 * vr + (1.f - beta2) * gv * gv;
    const float denom = sqrtf(vr) * inv_sqrt_bc2 + eps;
    pr = pr - step_size * (mr / denom);
    er = ema_beta * er + (1.f - ema_beta) * pr;
  };
  // 16-byte path: chunk starts are multiples of 2048 elements, so alignment is that of the tensors' base pointers
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)ema) & 15) == 0;
  long long iv = i0;
  if (vec) {
    const long long n4 = (i1 - i0) / 4;
    for (long long q = threadIdx.x; q < n4; q += blockDim.x) {
      const long long i = i0 + 4 * q;
      const float4 g4 = *reinterpret_cast<const float4*>(g + i);
      float4 m4 = *reinterpret_cast<const float4*>(m + i);
      float4 v4 = *reinterpret_cast<const float4*>(v + i);
      float4 p4 = *reinterpret_cast<const float4*>(p + i);
      float4 e4 = ema ? *reinterpret_cast<const float4*>(ema + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      update(g4.x, m4.x, v4.x, p4.x, e4.x);
      update(g4.y, m4.y, v4.y, p4.y, e4.y);
      update(g4.z, m4.z, v4.z, p4.z, e4.z);
      update(g4.w, m4.w, v4.w, p4.w, e4.w);
      *reinterpret_cast<float4*>(m + i) = m4;
      *reinterpret_cast<float4*>(v + i) = v4;
      *reinterpret_cast<float4*>(p + i) = p4;
      if (ema) *reinterpret_cast<float4*>(ema + i) = e4;
    }
    iv = i0 + 4 * n4;
  }
  for (long long i = iv + threadIdx.x; i < i1; i += blockDim.x) {
    float mr = m[i], vr = v[i], pr = p[i], er = ema ? ema[i] : 0.f;
    update(g[i], mr, vr, pr, er);
    m[i] = mr; v[i] = vr; p[i] = pr;
    if (ema) ema[i] = er;
  }
}

__global__ void bump_steps_kernel(const SgTensorEntry* __restrict__ tab, int T) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < T) *reinterpret_cast<float*>(tab[i].step) += 1.f;
}

// EMA only (parameters with no gradient this step still get averaged: models/__init__.py:33-36 loops over all)
__global__ void ema_multi_kernel(const SgTensorEntry* __restrict__ tab, int T, float ema_beta) {
  const int e = find_entry(tab, T, blockIdx.x);
  const SgTensorEntry en = tab[e];
  const float* p = reinterpret_cast<const float*>(en.p);
  float* ema = reinterpret_cast<float*>(en.ema);
  long long i0 = (long long)(blockIdx.x - en.chunk0) * CHUNK;
  long long i1 = i0 + CHUNK < en.n ? i0 + CHUNK : en.n;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) ema[i] = ema_beta * ema[i] + (1.f - ema_beta) * p[i];
}

// ---- Philox4x32-10 + Box-Muller ----------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__global__ void randn_kernel(float* __restrict__ out, long long n, unsigned long long seed, unsigned long long offset,
                             const unsigned long long* __restrict__ base_dev = nullptr) {
  if (base_dev) offset += *base_dev;                       // Philox counter base from device memory (CUDA-graph replay)
  const long long nq = (n + 3) / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
    unsigned long long ctr = offset + (unsigned long long)q;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x5347u, 0xB200u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    float u0 = ((float)c[0] + 0.5f) * (1.f / 4294967296.f);
    float u1 = ((float)c[1] + 0.5f) * (1.f / 4294967296.f);
    float u2 = ((float)c[2] + 0.5f) * (1.f / 4294967296.f);
    float u3 = ((float)c[3] + 0.5f) * (1.f / 4294967296.f);
    if (u0 < 1e-37f) u0 = 1e-37f;
    if (u2 < 1e-37f) u2 = 1e-37f;
    float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
    float s0, c0, s1, c1;
    sincospif(2.f * u1, &s0, &c0);
    sincospif(2.f * u3, &s1, &c1);
    float z[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (q * 4 + k < n) out[q * 4 + k] = z[k];
  }
}

}  // namespace

SG_EXPORT int sg_chunk_elems(void) { return CHUNK; }

SG_EXPORT int sg_sqnorm_multi(const SgTensorEntry* table_dev, int T, int total_chunks, double* out_dev, double* ws,
                              void* stream) {
  if (T <= 0 || total_chunks <= 0) return 0;
  SG_REQUIRE(ws != nullptr, "sg_sqnorm_multi: workspace of total_chunks doubles missing");
  sqnorm_multi_kernel<<<total_chunks, 256, 0, (cudaStream_t)stream>>>(table_dev, T, ws);
  sqnorm_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(ws, total_chunks, out_dev);
  return sg_check_launch("sg_sqnorm_multi");
}

SG_EXPORT int sg_adam_multi(const SgTensorEntry* table_dev, int T, int total_chunks, float lr, float beta1, float beta2,
                            float eps, float grad_scale, const double* normsq_dev, float max_norm, float ema_beta,
                            void* stream) {
  if (T <= 0 || total_chunks <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  adam_multi_kernel<<<total_chunks, 256, 0, st>>>(table_dev, T, lr, beta1, beta2, eps, grad_scale, normsq_dev, max_norm, ema_beta);
  int rc = sg_check_launch("sg_adam_multi");
  if (rc) return rc;
  bump_steps_kernel<<<sg_cdiv(T, 128), 128, 0, st>>>(table_dev, T);
  return sg_check_launch("sg_adam_multi/steps");
}

SG_EXPORT int sg_ema_multi(const SgTensorEntry* table_dev, int T, int total_chunks, float ema_beta, void* stream) {
  if (T <= 0 || total_chunks <= 0) return 0;
  ema_multi_kernel<<<total_chunks, 256, 0, (cudaStream_t)stream>>>(table_dev, T, ema_beta);
  return sg_check_launch("sg_ema_multi");
}

SG_EXPORT int sg_randn(float* out, long long n, unsigned long long seed, unsigned long long offset, void* stream) {
  randn_kernel<<<sg_grid_1d((n + 3) / 4, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset);
  return sg_check_launch("sg_randn");
}

// counter = *base (device memory) + delta: the same stream of numbers as sg_randn with offset = *base + delta
SG_EXPORT int sg_randn_dev(float* out, long long n, unsigned long long seed, const unsigned long long* base,
                           unsigned long long delta, void* stream) {
  SG_REQUIRE(base != nullptr, "sg_randn_dev: counter pointer missing");
  randn_kernel<<<sg_grid_1d((n + 3) / 4, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(out, n, seed, delta, base);
  return sg_check_launch("sg_randn_dev");
}
