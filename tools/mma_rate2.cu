// Micro-benchmark, part 2 (NOT yet run on hardware -- written for the next GPU session; the measured part 1 is
// tools/mma_rate.cu).  Questions it answers, all with operands resident in shared memory and one issuing warp:
//   mode 0  SS    cta_group::1, M=128        (baseline, must reproduce part 1: N=64 -> ~68 cycles)
//   mode 1  SS    cta_group::1, M=64         (the weight-gradient kernels' half-height accumulators)
//   mode 2  TS    A operand read from TMEM   (does the ~61-cycle floor disappear when A is not fetched from smem?)
//   mode 3  TS+cp one tcgen05.cp.128x256b (smem -> TMEM, one 128x16 bf16 A slice) issued before every TS MMA
//   mode 4  SS    cta_group::2, M=256        (CTA pair: each CTA supplies its 128 A rows and N/2 B rows)
// Every wait is time-bounded (trap after 4 s); run it under `timeout`.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/mma_rate2 tools/mma_rate2.cu -lcuda
#include "../stylegan.pytorch_b200/csrc/sg_tc_common.cuh"
#include <cooperative_groups.h>
#include <cstdio>
#include <vector>

namespace cg = cooperative_groups;
using namespace tcu;

constexpr int A_BYTES = 160 * 128;               // the product kernel's halo tile: 160 rows of 128 B
constexpr int B_BYTES = 9 * 128 * 128;           // nine taps of up to 128 rows of 128 B (mode 5)
constexpr int STAGES = 4;

struct __align__(8) Bars2 {
  uint64_t done;
  uint64_t scratch[STAGES];
  uint32_t tmem_base;
  uint32_t pad;
};

struct Cfg2 { int mode, M, N, groups; };

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t tmem_dst, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)0x3) : "memory");
}

__device__ __forceinline__ void fill_smem(uint8_t* smem, int bytes) {
  uint32_t* w = reinterpret_cast<uint32_t*>(smem);
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)(i + blockIdx.x * 7919) * 2654435761u;
    w[i] = 0x3F003F00u | (h & 0x807F807Fu);     // two bf16 of magnitude ~0.5..1
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- modes 0-3: one CTA per SM -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) mma_rate2_kernel(Cfg2 c, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  Bars2* bars = reinterpret_cast<Bars2*>(smem_b + B_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  fill_smem(smem, STAGES * A_BYTES + B_BYTES);
  if (threadIdx.x == 0) {
    mbar_init(&bars->done, 1);
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars->scratch[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(c.M, c.N);
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t b0 = make_desc_sw128(smem_u32(smem_b));
    const uint32_t tmem_a = tmem_base + 256;    // A slices (8 columns each) behind the accumulator columns
    const long long t0 = clock64();
    for (int g = 0; g < c.groups; ++g) {
      const int stage = g % STAGES;
      const uint64_t a0 = make_desc_sw128(smem_base + stage * A_BYTES);
      const bool leader = elect_one_sync();
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = a0 + (uint64_t)((kh * 2048 + k * 32) >> 4);
          const uint64_t bd = b0 + (uint64_t)((k * 32) >> 4);
          const uint32_t acc = (g | kh | k) ? 1u : 0u;
          if (leader) {
            if (c.mode <= 1) umma_bf16(tmem_base, ad, bd, idesc, acc);
            else {
              const uint32_t ta = tmem_a + (uint32_t)(((kh * 4 + k) & 7) * 8);
              if (c.mode == 3) tmem_cp_128x256b(ta, ad);
              umma_ts(tmem_base, ta, bd, idesc, acc);
            }
          }
        }
      }
      if (leader) umma_commit(&bars->scratch[stage]);
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(&bars->done);
    __syncwarp();
    mbar_wait(&bars->done, 0);
    if (lane == 0) cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- mode 4: CTA pair ----------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_rate_pair_kernel(Cfg2 c, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  Bars2* bars = reinterpret_cast<Bars2*>(smem_b + B_BYTES);
  cg::cluster_group cluster = cg::this_cluster();
  const bool is_leader_cta = cluster.block_rank() == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  fill_smem(smem, STAGES * A_BYTES + B_BYTES);
  if (threadIdx.x == 0) {
    mbar_init(&bars->done, 1);
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars->scratch[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                               // one warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster.sync();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  long long t0 = 0;
  if (warp == 1 && is_leader_cta) {
    const uint32_t idesc = make_idesc_bf16(256, c.N);
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t b0 = make_desc_sw128(smem_u32(smem_b));
    t0 = clock64();
    for (int g = 0; g < c.groups; ++g) {
      const int stage = g % STAGES;
      const uint64_t a0 = make_desc_sw128(smem_base + stage * A_BYTES);
      const bool leader = elect_one_sync();
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = a0 + (uint64_t)((kh * 2048 + k * 32) >> 4);
          // mode 5: a different B tile for every tap (kh, kw = g % 3), HB = N/2 rows of 128 B each, as the convolution has
          const uint32_t tap_off = c.mode == 5 ? (uint32_t)(((g % 3) * 3 + kh) * (c.N / 2) * 128) : 0u;
          const uint64_t bd = b0 + (uint64_t)((tap_off + k * 32) >> 4);
          // mode 6: alternate between two accumulators per group (as the convolution's tiles do)
          const uint32_t td = tmem_base + ((c.mode == 6 && ((g / 3) & 1)) ? (uint32_t)c.N : 0u);
          if (leader) umma_pair(td, ad, bd, idesc, (g | kh | k) ? 1u : 0u);
        }
      }
      if (leader) umma_commit_pair(&bars->scratch[stage]);
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit_pair(&bars->done);
    __syncwarp();
  }
  if (warp == 1) {                               // both CTAs receive the multicast commit
    mbar_wait(&bars->done, 0);
    if (lane == 0 && is_leader_cta) cycles[blockIdx.x >> 1] = clock64() - t0;
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* cyc;
  CK(cudaMalloc(&cyc, sms * sizeof(long long)));
  const size_t smem = STAGES * A_BYTES + B_BYTES + sizeof(Bars2) + 1024;
  CK(cudaFuncSetAttribute(mma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(mma_rate_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const char* names[] = {"SS M=128", "SS M=64", "TS (A in TMEM)", "TS + cp per MMA", "SS pair M=256", "pair, B per tap", "pair, 2 accum"};
  struct Row { int mode, M, N; };
  std::vector<Row> rows = {{0, 128, 64}, {0, 128, 128}, {0, 128, 256}, {1, 64, 64},  {1, 64, 128}, {1, 64, 192},
                           {1, 64, 256}, {2, 128, 64},  {2, 128, 128}, {2, 128, 256}, {3, 128, 64}, {3, 128, 128},
                           {4, 256, 64}, {4, 256, 128}, {4, 256, 256}, {5, 256, 64}, {5, 256, 128}, {5, 256, 256}, {6, 256, 64}, {6, 256, 128}};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<long long> h(sms);
  printf("%-18s %4s %4s | %10s | %8s %9s   (cycles per instruction; TFLOP/s from CUDA events over the whole chip)\n",
         "mode", "M", "N", "cyc/MMA", "ms", "TFLOP/s");
  for (const Row& r : rows) {
    Cfg2 c{r.mode, r.M, r.N, 600};
    auto launch = [&]() {
      if (r.mode >= 4) mma_rate_pair_kernel<<<sms / 2 * 2, 128, smem>>>(c, cyc);
      else mma_rate2_kernel<<<sms, 128, smem>>>(c, cyc);
    };
    launch();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 3;
    const double mmas = (double)c.groups * 12;
    const int ctas = r.mode >= 4 ? sms / 2 : sms;        // instructions are issued by one CTA per pair in mode 4
    CK(cudaMemcpy(h.data(), cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost));
    double cs = 0;
    for (int i = 0; i < ctas; ++i) cs += h[i] / mmas;
    const double cyc_per = cs / ctas;
    const double flops = mmas * ctas * 2.0 * r.M * r.N * 16;
    printf("%-18s %4d %4d | %10.1f | %8.4f %9.1f\n", names[r.mode], r.M, r.N, cyc_per, ms, flops / (ms * 1e-3) / 1e12);
    fflush(stdout);
  }
  return 0;
}
