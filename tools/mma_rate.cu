// Micro-benchmark: what paces tcgen05.mma (M=128, kind::f16, SS operands, K-major SW128) in the shape conv3x3_tc_kernel
// issues it?  One CTA per SM, the same 6-warp roles as the product kernel, each role switchable:
//   accmode 0: every MMA accumulates into one TMEM accumulator per tile (product behaviour)
//           1: consecutive MMAs alternate between two accumulators (two independent dependency chains)
//   issue   0: warp-uniform loop + elect.sync (product)      1: the whole loop inside `if (lane == 0)`
//   epi     1: four epilogue warps drain each finished accumulator (tcgen05.ld + bias/LeakyReLU + 256-bit stores)
//   tma     0: operands stay where they are in smem          2: full producer/consumer pipeline (TMA halo boxes from
//              a 32x256x256x64 tensor, MMA waits for the data, producer waits for the MMAs' commit)
// Prints cycles per MMA (clock64 inside the MMA warp, averaged over CTAs) and TFLOP/s from CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/mma_rate tools/mma_rate.cu -lcuda
#include "../stylegan.pytorch_b200/csrc/sg_tc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace tcu;

constexpr int TH = 8, TW = 16, BK = 64, STAGES = 6;
constexpr int A_BYTES = (TH + 2) * TW * BK * 2;  // 20480

struct __align__(8) Bars {
  uint64_t full[STAGES], empty[STAGES], tfull[2], tempty[2], done;
  uint32_t tmem_base;
  volatile uint32_t stop;
};

struct Cfg {
  int N, accmode, issue, epi, tma, tiles, gpt;   // gpt = groups (stages) per tile; each group = 12 MMAs
  int H, W, NB;
};

__global__ void fill_kernel(__nv_bfloat16* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u;
    h ^= h >> 15;
    p[i] = __float2bfloat16(((int)(h & 255) - 128) * (1.f / 128.f));
  }
}

__global__ void __launch_bounds__(192, 1)
mma_rate_kernel(const __grid_constant__ CUtensorMap map_x, Cfg c, long long* cycles, __nv_bfloat16* scratch) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  const int b_tile = c.N * BK * 2;
  Bars* bars = reinterpret_cast<Bars*>(smem_b + 2 * 256 * BK * 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // pseudo-random bf16 operands (generic-proxy writes -> visible to the async proxy after the fence)
  {
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    const int words = (STAGES * A_BYTES + 2 * 256 * BK * 2) / 4;
    for (int i = threadIdx.x; i < words; i += blockDim.x) {
      uint32_t h = (uint32_t)(i + blockIdx.x * 7919) * 2654435761u;
      // two bf16 in [-1, 1): sign/exponent of ~0.5, random mantissa
      w[i] = 0x3F003F00u | (h & 0x807F807Fu);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    mbar_init(&bars->tfull[0], 1); mbar_init(&bars->tfull[1], 1); mbar_init(&bars->done, 1);
    mbar_init(&bars->tempty[0], 4); mbar_init(&bars->tempty[1], 4);
    bars->stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&map_x);
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
  const int total_groups = c.tiles * c.gpt;

  if (warp == 0) {
    if (c.tma && lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const int tiles_w = c.W / TW, tiles_h = c.H / TH;
      for (int t = 0; t < c.tiles; ++t) {
        int r = (blockIdx.x + t * gridDim.x) % (c.NB * tiles_h * tiles_w);
        int wt = r % tiles_w; r /= tiles_w;
        int ht = r % tiles_h; int n = r / tiles_h;
        for (int g = 0; g < c.gpt; ++g) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          mbar_expect_tx(&bars->full[stage], A_BYTES);
          tma_load_4d(smem + stage * A_BYTES, &map_x, &bars->full[stage], 0, wt * TW + (g % 3) - 1, ht * TH - 1, n);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, c.N);
    const uint32_t smem_base = smem_u32(smem), sB0 = smem_u32(smem_b);
    long long t0 = clock64();
    auto body = [&](bool uniform) {
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < c.tiles; ++t) {
        const int acc = t & 1;
        if (c.epi) { mbar_wait(&bars->tempty[acc], ((t >> 1) & 1) ^ 1); tc_fence_after(); }
        for (int g = 0; g < c.gpt; ++g) {
          if (c.tma == 2) { mbar_wait(&bars->full[stage], phase); tc_fence_after(); }
          const uint64_t a0 = make_desc_sw128(smem_base + stage * A_BYTES);
          const uint64_t b0 = make_desc_sw128(sB0);
          const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32);
          const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
          const bool leader = uniform ? elect_one_sync() : true;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t al = a_lo + (uint32_t)((kh * (TW * 128) + k * 32) >> 4);
              const uint32_t bl = b_lo + (uint32_t)(((kh & 1) * b_tile + k * 32) >> 4);
              uint32_t d;
              if (c.accmode == 0) d = tmem_base + (uint32_t)(acc * c.N);
              else d = tmem_base + (uint32_t)((((kh * 4 + k) & 1)) * c.N);
              if (leader) umma_bf16_lohi(d, al, a_hi, bl, b_hi, idesc, (g | kh | (k >> (c.accmode ? 1 : 0))) ? 1u : 0u);
            }
          }
          if (leader) umma_commit(&bars->empty[stage]);
          if (uniform) __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (uniform) { if (elect_one_sync()) umma_commit(&bars->tfull[acc]); __syncwarp(); }
        else umma_commit(&bars->tfull[acc]);
      }
    };
    if (c.issue == 0) body(true);
    else if (lane == 0) body(false);
    __syncwarp();
    if (lane == 0) {
      umma_commit(&bars->done);
      mbar_wait(&bars->done, 0);
      long long t1 = clock64();
      cycles[blockIdx.x] = t1 - t0;
      bars->stop = 1;
    }
    __syncwarp();
  } else if (c.epi) {
    const int q = warp & 3, m = q * 32 + lane;
    uint32_t ph[2] = {0, 0};
    for (int t = 0; t < c.tiles; ++t) {
      const int acc = t & 1;
      mbar_wait(&bars->tfull[acc], ph[acc]);
      ph[acc] ^= 1;
      tc_fence_after();
      const int BNe = c.accmode == 0 ? c.N : (2 * c.N > 256 ? 256 : 2 * c.N);  // columns to drain
      __nv_bfloat16* dst = scratch + ((size_t)(blockIdx.x * 32 + (t & 31)) * 128 + m) * 256;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c.accmode == 0 ? acc * c.N : 0);
      for (int c0 = 0; c0 < BNe; c0 += 64) {
        uint32_t v[4][16];
        tmem_ld16_nowait(taddr + c0, v[0]);
        tmem_ld16_nowait(taddr + c0 + 16, v[1]);
        tmem_ld16_nowait(taddr + c0 + 32, v[2]);
        tmem_ld16_nowait(taddr + c0 + 48, v[3]);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t packed[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a = __uint_as_float(v[u][2 * j]) + 0.25f, b = __uint_as_float(v[u][2 * j + 1]) + 0.25f;
            a = a > 0.f ? a : 0.2f * a; b = b > 0.f ? b : 0.2f * b;
            __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
            packed[j] = *reinterpret_cast<uint32_t*>(&h2);
          }
          st_global_256(dst + c0 + u * 16, packed);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tempty[acc]);
    }
  }
  (void)total_groups;
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int NB = 32, H = 256, W = 256;
  __nv_bfloat16 *x, *scratch;
  long long* cyc;
  size_t xn = (size_t)NB * H * W * 64;
  CK(cudaMalloc(&x, xn * 2));
  CK(cudaMalloc(&scratch, (size_t)sms * 32 * 128 * 256 * 2));
  CK(cudaMalloc(&cyc, sms * sizeof(long long)));
  fill_kernel<<<1024, 256>>>(x, xn);
  CK(cudaDeviceSynchronize());
  CUtensorMap map;
  if (!encode_nhwc(&map, x, NB, H, W, 64, TW, TH + 2)) { printf("encode failed\n"); return 1; }
  size_t smem = STAGES * A_BYTES + 2 * 256 * BK * 2 + sizeof(Bars) + 1024;
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  printf("smem %zu B, %d SMs\n", smem, sms);
  printf("%4s %3s %5s %3s %3s %4s | %9s %9s %9s | %8s %8s\n", "N", "acc", "issue", "epi", "tma", "gpt", "cyc/MMA",
         "min", "max", "ms", "TFLOP/s");
  struct Row { int N, acc, issue, epi, tma, gpt; };
  std::vector<Row> rows = {
      {64, 0, 0, 0, 0, 3},  {64, 0, 1, 0, 0, 3},  {64, 1, 0, 0, 0, 3},  {64, 0, 0, 1, 0, 3},
      {64, 0, 0, 0, 2, 3},  {64, 0, 0, 1, 2, 3},  {64, 1, 0, 0, 2, 3},  {64, 0, 0, 1, 2, 6},
      {128, 0, 0, 0, 0, 3}, {128, 0, 1, 0, 0, 3}, {128, 1, 0, 0, 0, 3}, {128, 0, 0, 1, 2, 3}, {128, 0, 0, 1, 2, 12},
      {256, 0, 0, 0, 0, 3}, {256, 0, 0, 1, 2, 3}, {256, 0, 0, 1, 2, 12},
      {32, 0, 0, 0, 0, 3},  {16, 0, 0, 0, 0, 3},
  };
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<long long> h(sms);
  for (const Row& r : rows) {
    Cfg c;
    c.N = r.N; c.accmode = r.acc; c.issue = r.issue; c.epi = r.epi; c.tma = r.tma; c.gpt = r.gpt;
    c.tiles = 111 * 3 / r.gpt * 2;   // ~2x the product kernel's per-CTA work at the north-star shape
    if (c.tiles < 8) c.tiles = 8;
    c.H = H; c.W = W; c.NB = NB;
    mma_rate_kernel<<<sms, 192, smem>>>(map, c, cyc, scratch);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < 3; ++i) mma_rate_kernel<<<sms, 192, smem>>>(map, c, cyc, scratch);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 3;
    CK(cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    const double mmas = (double)c.tiles * c.gpt * 12;
    double sum = 0, mn = 1e30, mx = 0;
    for (long long v : h) { double d = v / mmas; sum += d; mn = d < mn ? d : mn; mx = d > mx ? d : mx; }
    const double flops = mmas * sms * 2.0 * 128 * r.N * 16;
    printf("%4d %3d %5d %3d %3d %4d | %9.1f %9.1f %9.1f | %8.4f %8.1f\n", r.N, r.acc, r.issue, r.epi, r.tma, r.gpt,
           sum / sms, mn, mx, ms, flops / (ms * 1e-3) / 1e12);
    fflush(stdout);
  }
  return 0;
}
