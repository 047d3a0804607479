// Shared helpers for the sgb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define SG_EXPORT extern "C" __attribute__((visibility("default")))

// dtype codes of the C-ABI (include/sgb200.h)
#define SG_F32 0
#define SG_BF16 1

#define SG_LRELU_SLOPE 0.2f

void sg_set_error(const char* msg);
int sg_check_launch(const char* what);

#define SG_REQUIRE(cond, msg)            \
  do {                                   \
    if (!(cond)) {                       \
      sg_set_error(msg);                 \
      return 1000;                       \
    }                                    \
  } while (0)

// Dispatch a templated launcher on the dtype code.
#define SG_DISPATCH(dt, ...)                                   \
  do {                                                         \
    if ((dt) == SG_F32) { using T = float; __VA_ARGS__; }      \
    else if ((dt) == SG_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { sg_set_error("bad dtype code"); return 1001; }      \
  } while (0)

__device__ __forceinline__ float sg_to_f(float v) { return v; }
__device__ __forceinline__ float sg_to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <class T> __device__ __forceinline__ T sg_from_f(float v);
template <> __device__ __forceinline__ float sg_from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 sg_from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float sg_lrelu(float v) { return v > 0.f ? v : SG_LRELU_SLOPE * v; }

// 16-byte vector of T: 4 floats or 8 bf16
template <class T> struct SgVec;
template <> struct SgVec<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ void zero() { raw = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ float get(int i) const { return (&raw.x)[i]; }
  __device__ __forceinline__ void set(int i, float v) { (&raw.x)[i] = v; }
};
template <> struct SgVec<__nv_bfloat16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void zero() { raw = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ float get(int i) const {
    uint32_t w = (&raw.x)[i >> 1];
    return __uint_as_float((i & 1) ? (w & 0xffff0000u) : (w << 16));
  }
  __device__ __forceinline__ void set(int i, float v) {
    uint32_t b = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
    uint32_t& w = (&raw.x)[i >> 1];
    w = (i & 1) ? ((w & 0x0000ffffu) | (b << 16)) : ((w & 0xffff0000u) | b);
  }
};

__device__ __forceinline__ float sg_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum (blockDim.x <= 1024, multiple of 32).  Result valid in thread 0.
__device__ __forceinline__ float sg_block_sum(float v, float* smem32) {
  v = sg_warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? smem32[lane] : 0.f;
    v = sg_warp_sum(v);
  }
  return v;
}

static inline int sg_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline int sg_grid_1d(long long work, int block, int cap = 148 * 16) {
  long long g = (work + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}
