// tcgen05 / TMA weight gradient of the 3x3 stride-1 convolution (bf16 operands, fp32 accumulation in TMEM).
//
//   part[s][kh*3+kw][co][ci] = sum_{pixels p of split s} dy[p][co] * x[p + (kh-1, kw-1)][ci]
//
// reference: the autograd adjoint of F.conv2d at models/CustomLayers.py:170-171 wrt the weight (and, with the
// operands swapped, the weight gradient that the R1 double backward needs).
//
// GEMM view: M = Cout block (128, or 64 -> half-datapath accumulator layout), N = Cin block (<= 128), K = pixels.
// Both operands are read straight from the NHWC activations as MN-major 128B-swizzled tiles (a pixel row of 64
// channels is one 128-byte swizzle row), so no transpose pass exists anywhere.  One CTA owns (kw, M block, N block,
// pixel split): per 8x16-pixel tile it TMA-loads the dy tile and the kw-shifted (TH+2)-row halo box of x, and the
// three vertical taps are the same halo buffer entered 0/16/32 rows down -> three accumulators of BN columns in TMEM.
// Pixel splits are reduced (and scaled, and written in OIHW) by sg_wprep_bwd.
#include "sg_common.cuh"
#include "../../include/sgb200.h"
#include "sg_tc_common.cuh"

namespace wg {

using namespace tcu;

constexpr int TH = 8, TW = 16;
constexpr int DY_BLOCK_BYTES = TH * TW * 128;          // 16384: 128 pixels x 64 channels
constexpr int X_BLOCK_BYTES = (TH + 2) * TW * 128;     // 20480: 160 halo pixels x 64 channels
constexpr int NUM_THREADS = 192;
constexpr int MAX_STAGES = 8;

struct __align__(8) Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t done;
  uint32_t tmem_base;
  uint32_t pad;
};

struct Params {
  int N, H, W, Cin, Cout, BM, BN;
  int tiles_h, tiles_w, total_tiles;
  int MB, NB, S;
  int stages, stage_bytes, tmem_cols;
  float* part;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad3x3_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + (size_t)p.stages * p.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work unit
  int u = blockIdx.x;
  const int nb = u % p.NB; u /= p.NB;
  const int mb = u % p.MB; u /= p.MB;
  const int kw = u % 3;
  const int s = u / 3;
  const int per = (p.total_tiles + p.S - 1) / p.S;
  const int t0 = s * per;
  const int t1 = (t0 + per < p.total_tiles) ? (t0 + per) : p.total_tiles;
  const int a_blocks = p.BM / 64, b_blocks = p.BN / 64;
  const int a_bytes = a_blocks * DY_BLOCK_BYTES;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    mbar_init(&bars->done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&map_dy);
    prefetch_tmap(&map_x);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t0; tile < t1; ++tile) {
        int r = tile;
        const int wt = r % p.tiles_w; r /= p.tiles_w;
        const int ht = r % p.tiles_h;
        const int n = r / p.tiles_h;
        const int h0 = ht * TH, w0 = wt * TW;
        mbar_wait(&bars->empty[stage], phase ^ 1);
        uint8_t* sA = smem + (size_t)stage * p.stage_bytes;
        uint8_t* sB = sA + a_bytes;
        mbar_expect_tx(&bars->full[stage], (uint32_t)(a_bytes + b_blocks * X_BLOCK_BYTES));
        for (int j = 0; j < a_blocks; ++j)
          tma_load_4d(sA + j * DY_BLOCK_BYTES, &map_dy, &bars->full[stage], mb * p.BM + 64 * j, w0, h0, n);
        for (int j = 0; j < b_blocks; ++j)
          tma_load_4d(sB + j * X_BLOCK_BYTES, &map_x, &bars->full[stage], nb * p.BN + 64 * j, w0 + kw - 1, h0 - 1, n);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // whole warp, warp-uniform state; one elected lane issues (see sg_conv_tc.cu)
    const uint32_t idesc = make_idesc_bf16_mn(p.BM, p.BN);
    const uint32_t smem_base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = t0; tile < t1; ++tile) {
      mbar_wait(&bars->full[stage], phase);
      tc_fence_after();
      const uint32_t sA = smem_base + (uint32_t)(stage * p.stage_bytes);
      const uint32_t sB = sA + (uint32_t)a_bytes;
      const uint64_t a0 = make_desc_mn_sw128(sA, DY_BLOCK_BYTES);
      const uint64_t b0 = make_desc_mn_sw128(sB, X_BLOCK_BYTES);
      const uint32_t not_first = (tile != t0) ? 1u : 0u;
      const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32);
      const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
      const bool leader = elect_one_sync();
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int k = 0; k < (TH * TW) / 16; ++k) {            // 8 K-steps of 16 pixels = 2048 B each
          const uint32_t al = a_lo + (uint32_t)((k * 2048) >> 4);
          const uint32_t bl = b_lo + (uint32_t)((kh * (TW * 128) + k * 2048) >> 4);
          if (leader) umma_bf16_lohi(tmem_base + (uint32_t)(kh * p.BN), al, a_hi, bl, b_hi, idesc, k ? 1u : not_first);
        }
      }
      if (leader) umma_commit(&bars->empty[stage]);
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one_sync()) umma_commit(&bars->done);
    __syncwarp();
  } else {
    // epilogue: TMEM -> fp32 partials.  M=128: row = 32*q + lane.  M=64: row = 16*q + lane for lane < 16.
    const int q = warp & 3;
    mbar_wait(&bars->done, 0);
    tc_fence_after();
    const bool m64 = (p.BM == 64);
    const int row = m64 ? (q * 16 + lane) : (q * 32 + lane);
    const bool active = m64 ? (lane < 16) : true;
    const int co = mb * p.BM + row;
    for (int kh = 0; kh < 3; ++kh) {
      float* dst = p.part + ((((size_t)s * 9 + (kh * 3 + kw)) * p.Cout + co) * p.Cin + (size_t)nb * p.BN);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kh * p.BN);
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        if (active && co < p.Cout && nb * p.BN + c0 < p.Cin) {
          float4* d4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            d4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                __uint_as_float(v[4 * j + 3]));
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// Channel block of the GEMM's M / N dimension.  Channel counts below 64 use a 64-wide block: the TMA box reads past
// the tensor's channel extent, which arrives as zeros, and the epilogue writes only the real rows / columns.
static int pick_block(int C) {
  if (C % 128 == 0) return 128;
  if (C == 64) return 64;
  if (C < 64 && C % 16 == 0) return 64;
  return 0;
}

}  // namespace wg

SG_EXPORT int sg_wgrad3x3_tc_supported(int N, int H, int W, int Cin, int Cout) {
  if (N < 1 || H % wg::TH != 0 || W % wg::TW != 0) return 0;
  if (wg::pick_block(Cin) == 0 || wg::pick_block(Cout) == 0) return 0;
  return 1;
}

// Number of pixel splits the kernel will use (the caller sizes `part` as [S][9][Cout][Cin] floats).
SG_EXPORT int sg_wgrad3x3_tc_splits(int N, int H, int W, int Cin, int Cout) {
  if (!sg_wgrad3x3_tc_supported(N, H, W, Cin, Cout)) return 0;
  int total = N * (H / wg::TH) * (W / wg::TW);
  int units = 3 * ((Cout + wg::pick_block(Cout) - 1) / wg::pick_block(Cout)) * ((Cin + wg::pick_block(Cin) - 1) / wg::pick_block(Cin));
  int S = (148 + units - 1) / units;
  if (S > total) S = total;
  if (S < 1) S = 1;
  int per = (total + S - 1) / S;
  S = (total + per - 1) / per;            // no empty split: every CTA initialises its accumulators
  return S;
}

SG_EXPORT int sg_wgrad3x3_tc(const void* x, const void* dy, float* part, int N, int H, int W, int Cin, int Cout,
                             void* stream) {
  using namespace wg;
  SG_REQUIRE(sg_wgrad3x3_tc_supported(N, H, W, Cin, Cout), "sg_wgrad3x3_tc: shape not supported");
  CUtensorMap map_dy, map_x;
  SG_REQUIRE(encode_nhwc(&map_dy, dy, N, H, W, Cout, TW, TH), "sg_wgrad3x3_tc: cuTensorMapEncodeTiled(dy) failed");
  SG_REQUIRE(encode_nhwc(&map_x, x, N, H, W, Cin, TW, TH + 2), "sg_wgrad3x3_tc: cuTensorMapEncodeTiled(x) failed");
  Params p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.BM = pick_block(Cout); p.BN = pick_block(Cin);
  p.tiles_h = H / TH; p.tiles_w = W / TW; p.total_tiles = N * p.tiles_h * p.tiles_w;
  p.MB = (Cout + p.BM - 1) / p.BM; p.NB = (Cin + p.BN - 1) / p.BN;
  p.S = sg_wgrad3x3_tc_splits(N, H, W, Cin, Cout);
  p.stage_bytes = (p.BM / 64) * DY_BLOCK_BYTES + (p.BN / 64) * X_BLOCK_BYTES;
  int stages = (216 * 1024) / p.stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  SG_REQUIRE(stages >= 2, "sg_wgrad3x3_tc: not enough shared memory");
  p.stages = stages;
  int cols = 3 * p.BN, tc_cols = 32;
  while (tc_cols < cols) tc_cols <<= 1;
  p.tmem_cols = tc_cols;
  p.part = part;
  size_t smem = (size_t)stages * p.stage_bytes + sizeof(Barriers) + 1024;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(wgrad3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  SG_REQUIRE(attr_err == cudaSuccess, "sg_wgrad3x3_tc: cannot raise the dynamic shared memory limit");
  int grid = p.S * 3 * p.MB * p.NB;
  wgrad3x3_tc_kernel<<<grid, NUM_THREADS, smem, (cudaStream_t)stream>>>(map_dy, map_x, p);
  return sg_check_launch("sg_wgrad3x3_tc");
}
