// tcgen05 / TMA weight gradient of the 3x3 stride-1 convolution, second generation: every MMA is 128 x 192 x 16.
//
//   part[s][kh*3+kw][co][ci] = sum_{pixels p of split s} dy[p][co] * x[p + (kh-1, kw-1)][ci]
//
// reference: the autograd adjoint of F.conv2d at models/CustomLayers.py:170-171 wrt the weight (and, with the operands
// swapped, the weight gradient the R1 double backward needs, models/Losses.py:197-210).
//
// The first kernel (sg_wgrad_tc.cu) issues 64/128 x 64/128 x 16 MMAs, one accumulator per vertical tap: at 64 channels
// that is a half-datapath M = 64 instruction with N = 64, 23 % of the tensor peak (421 TFLOP/s at 64->64 @256x256).
// Here the GEMM N dimension is the THREE vertical taps of one 64-channel block of x: the (TH+2)-row halo box is an
// MN-major operand whose 64-element blocks are 2048 bytes (one 16-pixel image row) apart, so blocks 0/1/2 of the same
// buffer are the taps kh = 0/1/2 -> N = 192 in one instruction.  The M dimension is 128:
//   mode 1 (Cout % 128 == 0): two 64-channel blocks of dy; one CTA per (kw, 128 output channels, 1-2 x blocks, split);
//   mode 0 (Cout == 64):      dy shifted by kw = 0 and kw = 1 stacked (two TMA boxes), plus an M = 64 instruction for
//                             kw = 2: one CTA computes all nine taps of a split from ONE x halo box.
// (dW[kh,kw] = sum_q dy[q - (0,kw-1)] x[q + (kh-1,0)]: shifting dy instead of x keeps the x box 16-pixel aligned.)
#include "sg_common.cuh"
#include "../../include/sgb200.h"
#include "sg_tc_common.cuh"

namespace wg2 {

using namespace tcu;

constexpr int TH = 8, TW = 16;
constexpr int DY_BLOCK_BYTES = TH * TW * 128;          // 16384: 128 pixels x 64 channels
constexpr int X_BLOCK_BYTES = (TH + 2) * TW * 128;     // 20480: 160 halo pixels x 64 channels
constexpr int ROW_BYTES = TW * 128;                    // 2048: one 16-pixel image row = one K step = one kh block
constexpr int NUM_THREADS = 192;
constexpr int MAX_STAGES = 4;

struct __align__(8) Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t done;
  uint32_t tmem_base;
  uint32_t pad;
};

struct Params {
  int N, H, W, Cin, Cout, mode;
  int tiles_h, tiles_w, total_tiles;
  int MB, NBG, XB;                                     // 128-channel dy blocks, x block groups, x blocks per group
  int S, stages, stage_bytes, a_bytes, tmem_cols;
  float* part;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad3x3_v2_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + (size_t)p.stages * p.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int u = blockIdx.x;
  const int nbg = u % p.NBG; u /= p.NBG;
  int mb = 0, kw = 0;
  if (p.mode == 1) { mb = u % p.MB; u /= p.MB; kw = u % 3; u /= 3; }
  const int s = u;
  const int per = (p.total_tiles + p.S - 1) / p.S;
  const int t0 = s * per;
  const int t1 = (t0 + per < p.total_tiles) ? (t0 + per) : p.total_tiles;

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
    // ===================== TMA producer =====================
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
        uint8_t* sB = sA + p.a_bytes;
        mbar_expect_tx(&bars->full[stage], (uint32_t)(p.a_bytes + p.XB * X_BLOCK_BYTES));
        if (p.mode == 0) {
          for (int j = 0; j < 3; ++j)                     // dy shifted by -(kw-1) columns, kw = j (zero outside the image)
            tma_load_4d(sA + j * DY_BLOCK_BYTES, &map_dy, &bars->full[stage], 0, w0 + 1 - j, h0, n);
          tma_load_4d(sB, &map_x, &bars->full[stage], nbg * 64, w0, h0 - 1, n);
        } else {
          for (int j = 0; j < 2; ++j)
            tma_load_4d(sA + j * DY_BLOCK_BYTES, &map_dy, &bars->full[stage], mb * 128 + 64 * j, w0, h0, n);
          for (int j = 0; j < p.XB; ++j)
            tma_load_4d(sB + j * X_BLOCK_BYTES, &map_x, &bars->full[stage], (nbg * p.XB + j) * 64, w0 + kw - 1, h0 - 1, n);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: whole warp, warp-uniform state, one elected lane issues =====================
    const uint32_t idesc128 = make_idesc_bf16_mn(128, 192);
    const uint32_t idesc64 = make_idesc_bf16_mn(64, 192);
    const uint32_t smem_base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = t0; tile < t1; ++tile) {
      mbar_wait(&bars->full[stage], phase);
      tc_fence_after();
      const uint32_t sA = smem_base + (uint32_t)(stage * p.stage_bytes);
      const uint32_t sB = sA + (uint32_t)p.a_bytes;
      const uint64_t a0 = make_desc_mn_sw128(sA, DY_BLOCK_BYTES);
      const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32);
      const uint32_t not_first = (tile != t0) ? 1u : 0u;
      const bool leader = elect_one_sync();
      if (p.mode == 0) {
        const uint64_t b0 = make_desc_mn_sw128(sB, ROW_BYTES);
        const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
        const uint32_t a2_lo = a_lo + (uint32_t)((2 * DY_BLOCK_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < (TH * TW) / 16; ++k) {
          const uint32_t ko = (uint32_t)((k * ROW_BYTES) >> 4);
          if (leader) {
            umma_bf16_lohi(tmem_base, a_lo + ko, a_hi, b_lo + ko, b_hi, idesc128, k ? 1u : not_first);
            umma_bf16_lohi(tmem_base + 192u, a2_lo + ko, a_hi, b_lo + ko, b_hi, idesc64, k ? 1u : not_first);
          }
        }
      } else {
        for (int xb = 0; xb < p.XB; ++xb) {
          const uint64_t b0 = make_desc_mn_sw128(sB + (uint32_t)(xb * X_BLOCK_BYTES), ROW_BYTES);
          const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
#pragma unroll
          for (int k = 0; k < (TH * TW) / 16; ++k) {
            const uint32_t ko = (uint32_t)((k * ROW_BYTES) >> 4);
            if (leader)
              umma_bf16_lohi(tmem_base + (uint32_t)(xb * 192), a_lo + ko, a_hi, b_lo + ko, b_hi, idesc128,
                             k ? 1u : not_first);
          }
        }
      }
      if (leader) umma_commit(&bars->empty[stage]);
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one_sync()) umma_commit(&bars->done);
    __syncwarp();
  } else {
    // ===================== epilogue: TMEM -> fp32 partials =====================
    const int q = warp & 3;
    mbar_wait(&bars->done, 0);
    tc_fence_after();
    auto drain = [&](uint32_t col0, int co, int kwi, int ci0, bool active) {
      // 192 columns = kh-major: column kh*64 + c is tap (kh, kwi), input channel ci0 + c
      for (int kh = 0; kh < 3; ++kh) {
        float* dst = p.part + ((((size_t)s * 9 + (kh * 3 + kwi)) * p.Cout + co) * p.Cin + ci0);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + col0 + (uint32_t)(kh * 64);
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          if (active) {
            float4* d4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              d4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                  __uint_as_float(v[4 * j + 3]));
          }
        }
      }
    };
    if (p.mode == 0) {
      const int row = q * 32 + lane;                       // M = 128: rows 0-63 = kw 0, rows 64-127 = kw 1
      drain(0u, row & 63, row >> 6, nbg * 64, true);
      drain(192u, q * 16 + lane, 2, nbg * 64, lane < 16);   // M = 64: row 16q+i lives in lane 32q+i
    } else {
      const int row = q * 32 + lane;
      for (int xb = 0; xb < p.XB; ++xb) drain((uint32_t)(xb * 192), mb * 128 + row, kw, (nbg * p.XB + xb) * 64, true);
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

static bool plan(int N, int H, int W, int Cin, int Cout, Params* p) {
  if (N < 1 || H % TH != 0 || W % TW != 0 || Cin % 64 != 0 || Cin < 64) return false;
  if (Cout == 64) {
    p->mode = 0; p->MB = 1; p->XB = 1; p->NBG = Cin / 64;
    p->a_bytes = 3 * DY_BLOCK_BYTES;
  } else if (Cout % 128 == 0) {
    p->mode = 1; p->MB = Cout / 128;
    p->XB = (Cin % 128 == 0) ? 2 : 1;
    p->NBG = Cin / (64 * p->XB);
    p->a_bytes = 2 * DY_BLOCK_BYTES;
  } else {
    return false;
  }
  p->N = N; p->H = H; p->W = W; p->Cin = Cin; p->Cout = Cout;
  p->tiles_h = H / TH; p->tiles_w = W / TW; p->total_tiles = N * p->tiles_h * p->tiles_w;
  const int units = (p->mode == 0 ? 1 : 3 * p->MB) * p->NBG;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int S = sms / units;                                   // one wave of CTAs
  if (S < 1) S = 1;
  if (S > p->total_tiles) S = p->total_tiles;
  const int per = (p->total_tiles + S - 1) / S;
  p->S = (p->total_tiles + per - 1) / per;               // no empty split: every CTA initialises its accumulators
  p->stage_bytes = p->a_bytes + p->XB * X_BLOCK_BYTES;
  int stages = (216 * 1024) / p->stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return false;
  p->stages = stages;
  p->tmem_cols = 512;                                    // 2 x 192 columns in either mode -> next power of two
  return true;
}

}  // namespace wg2

SG_EXPORT int sg_wgrad3x3_v2_supported(int N, int H, int W, int Cin, int Cout) {
  wg2::Params p;
  return wg2::plan(N, H, W, Cin, Cout, &p) ? 1 : 0;
}

// pixel splits (the caller sizes `part` as [S][9][Cout][Cin] floats), 0 if unsupported
SG_EXPORT int sg_wgrad3x3_v2_splits(int N, int H, int W, int Cin, int Cout) {
  wg2::Params p;
  return wg2::plan(N, H, W, Cin, Cout, &p) ? p.S : 0;
}

SG_EXPORT int sg_wgrad3x3_v2(const void* x, const void* dy, float* part, int N, int H, int W, int Cin, int Cout,
                             void* stream) {
  using namespace wg2;
  Params p;
  SG_REQUIRE(plan(N, H, W, Cin, Cout, &p), "sg_wgrad3x3_v2: shape not supported");
  CUtensorMap map_dy, map_x;
  SG_REQUIRE(encode_nhwc(&map_dy, dy, N, H, W, Cout, TW, TH), "sg_wgrad3x3_v2: cuTensorMapEncodeTiled(dy) failed");
  SG_REQUIRE(encode_nhwc(&map_x, x, N, H, W, Cin, TW, TH + 2), "sg_wgrad3x3_v2: cuTensorMapEncodeTiled(x) failed");
  p.part = part;
  const size_t smem = (size_t)p.stages * p.stage_bytes + sizeof(Barriers) + 1024;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(wgrad3x3_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  SG_REQUIRE(attr_err == cudaSuccess, "sg_wgrad3x3_v2: cannot raise the dynamic shared memory limit");
  const int grid = p.S * (p.mode == 0 ? 1 : 3 * p.MB) * p.NBG;
  wgrad3x3_v2_kernel<<<grid, NUM_THREADS, smem, (cudaStream_t)stream>>>(map_dy, map_x, p);
  return sg_check_launch("sg_wgrad3x3_v2");
}
