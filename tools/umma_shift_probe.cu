// Probe: how does a K-major SWIZZLE_128B UMMA operand descriptor behave when its start address is NOT a multiple of the
// 1024-byte swizzle atom, and when the stride between 8-row groups (SBO) is not a multiple of 1024?
//
// Why: the 3x3 convolution loads three horizontally shifted TMA boxes of the same activation rows (kw = 0,1,2), 60 KB per
// 64-channel tile, only to keep every descriptor 1024-byte aligned.  If the tensor core applies the 128B swizzle to the
// absolute shared-memory address (as TMA does when it writes), ONE halo box (18 rows x 10 px, rows 1280 B apart) can
// serve all nine taps: tap (kh, kw) = start address + (kh*10 + kw)*128, SBO = 1280.
//
// Method: A rows are written by software in the TMA layout (row R of 128 B at base + 128 R, 16-byte chunk c at position
// c ^ (R & 7)), B = 64 x 64 identity, so D[m][n] = A[row(m)][n].  Pass 0 stores the row number in every element, pass 1
// the column number; D tells which row and which chunk each (m, n) was read from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/umma_shift_probe tools/umma_shift_probe.cu -lcuda
#include "../stylegan.pytorch_b200/csrc/sg_tc_common.cuh"
#include <cstdio>
#include <vector>

using namespace tcu;

constexpr int A_ROWS = 320;                        // rows of 128 B available to the A operand
struct __align__(8) PBars { uint64_t done; uint32_t tmem_base; uint32_t pad; };
struct PCfg { int shift_rows, sbo, base_off, pass; };

__device__ __forceinline__ uint64_t desc_sw128_ex(uint32_t addr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7u) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(PCfg c, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_ROWS * 128;
  PBars* bars = reinterpret_cast<PBars*>(sB + 64 * 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < A_ROWS * 64; i += blockDim.x) {
    const int R = i >> 6, k = i & 63;
    const float v = c.pass == 0 ? (float)(R & 255) : (float)k;
    *reinterpret_cast<__nv_bfloat16*>(sA + R * 128 + (((k >> 3) ^ (R & 7)) << 4) + (k & 7) * 2) = __float2bfloat16(v);
  }
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int n = i >> 6, k = i & 63;
    *reinterpret_cast<__nv_bfloat16*>(sB + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) =
        __float2bfloat16(n == k ? 1.f : 0.f);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    mbar_init(&bars->done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, 64);
    const uint64_t a0 = desc_sw128_ex(smem_u32(sA) + (uint32_t)c.shift_rows * 128u, (uint32_t)c.sbo, (uint32_t)c.base_off);
    const uint64_t b0 = desc_sw128_ex(smem_u32(sB), 1024u, 0u);
    if (elect_one_sync()) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, a0 + (uint64_t)((k * 32) >> 4), b0 + (uint64_t)((k * 32) >> 4), idesc, k ? 1u : 0u);
      umma_commit(&bars->done);
    }
    __syncwarp();
  }
  mbar_wait(&bars->done, 0);
  tc_fence_after();
  {
    const int m = warp * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) out[m * 64 + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

int main() {
  const size_t smem = A_ROWS * 128 + 64 * 128 + sizeof(PBars) + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  float* d_out;
  cudaMalloc(&d_out, 128 * 64 * sizeof(float));
  std::vector<float> h(128 * 64);
  const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 10, 11, 12, 21, 22};
  const int sbos[] = {1024, 1280, 2048, 2304};
  printf("%-6s %-6s %-4s | %-22s %-22s\n", "shift", "SBO", "bo", "rows (pass 0)", "chunks (pass 1)");
  for (int sbo : sbos)
    for (int s : shifts)
      for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
        const int bo = bo_mode ? (s & 7) : 0;
        if (bo_mode && bo == 0) continue;
        int bad[2] = {0, 0};
        int first_m[2] = {-1, -1}, first_n[2] = {-1, -1};
        float first_v[2] = {0, 0}, first_e[2] = {0, 0};
        for (int pass = 0; pass < 2; ++pass) {
          PCfg c{s, sbo, bo, pass};
          probe_kernel<<<1, 128, smem>>>(c, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error %s at shift %d sbo %d bo %d\n", cudaGetErrorString(e), s, sbo, bo); return 1; }
          cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
              const int R = s + (m / 8) * (sbo / 128) + (m % 8);
              const float want = pass == 0 ? (float)(R & 255) : (float)n;
              if (h[m * 64 + n] != want) {
                if (!bad[pass]) { first_m[pass] = m; first_n[pass] = n; first_v[pass] = h[m * 64 + n]; first_e[pass] = want; }
                ++bad[pass];
              }
            }
        }
        char r0[64], r1[64];
        if (bad[0]) snprintf(r0, sizeof r0, "BAD %d (m%d n%d %g!=%g)", bad[0], first_m[0], first_n[0], first_v[0], first_e[0]);
        else snprintf(r0, sizeof r0, "ok");
        if (bad[1]) snprintf(r1, sizeof r1, "BAD %d (m%d n%d %g!=%g)", bad[1], first_m[1], first_n[1], first_v[1], first_e[1]);
        else snprintf(r1, sizeof r1, "ok");
        printf("%-6d %-6d %-4d | %-22s %-22s\n", s, sbo, bo, r0, r1);
      }
  cudaFree(d_out);
  return 0;
}
