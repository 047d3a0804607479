// tcgen05 / TMA up- and down-sampling convolutions in their tap-merged low-resolution form (bf16, fp32 TMEM).
//
//   UP   (4 output phases x 2x2 taps):  y[n, 2i+py, 2j+px, :] = sum_{a,b} x[n, i+a+py-1, j+b+px-1, :] . E[ph][a*2+b]
//        = nearest x2 -> conv3x3 (models/CustomLayers.py:153-154,172-173) or conv_transpose2d of the box-summed kernel
//          (:143-152), and the dgrad of the down convolution;
//   DOWN (4x4 taps, stride 2, pad 1):   y[n, i, j, :] = sum_{u,v} x[n, 2i-1+u, 2j-1+v, :] . E[u*4+v]
//        = conv3x3 -> 2x2 box (:156-180), and the dgrad of the up convolution.
// The effective weights E[tap][Cout][Cin] come from sg_wprep_fwd with the host tables of models/_geom.py, so this
// kernel does 16 tap-GEMMs per low-resolution pixel instead of the 36 of the unfused formulation and touches neither
// an up-sampled copy of the input nor a full-resolution copy of the output.
//
// Same structure as sg_conv_tc.cu: M = 8x16 tile of the low-resolution grid, one TMA box of (TH+1) rows per horizontal
// tap position (for DOWN the box is fetched with element strides 2 in W and H, one box per row parity), the two
// vertical taps are UMMA descriptors 0/16 rows into the box; weights stream with each stage.
#include "sg_common.cuh"
#include "../../include/sgb200.h"
#include "sg_tc_common.cuh"

int g_sg_pair_updown = 2;      // sg_fast_paths-style switch for A/B tests (sg_pair_updown): 0 single CTA, 1 first pair kernel, 2 halo-box pair kernel

int sg_convud_halo_supported(int mode, int N, int H, int W, int Cin, int Cout);      // csrc/sg_convud2_tc.cu
int sg_convud_halo(int mode, const void* x, const void* E, void* y, const float* bias, float bias_mul, int act, int N, int H,
                   int W, int Cin, int Cout, void* stream);

namespace ud {

using namespace tcu;

constexpr int TH = 8, TW = 16, BK = 64;
constexpr int A_BYTES = (TH + 1) * TW * BK * 2;   // 18432 = 18 KB (1024-aligned)
constexpr int NUM_THREADS = 192;
constexpr int MAX_STAGES = 8;
constexpr int MODE_UP = 0, MODE_DOWN = 1;

struct __align__(8) Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

struct Params {
  int mode, N, Hq, Wq, Cin, Cout, BN, Hout, Wout;
  int tiles_h, tiles_w, tiles_n, num_spatial, nphase, nbox;
  int kchunks, stages, stage_bytes, tmem_cols;
  const float* bias;
  float bias_mul;
  int act;
  __nv_bfloat16* y;
};

// box -> (TMA coordinate offsets, weight-tile indices of its two vertical taps)
struct BoxInfo { int dw, dh, w0, w1; };
__device__ __forceinline__ BoxInfo box_info(int mode, int phase, int box) {
  BoxInfo bi;
  if (mode == MODE_UP) {
    const int py = phase >> 1, px = phase & 1, b = box;       // taps a = 0,1 are rows 0 / TW of the box
    bi.dw = b + px - 1;
    bi.dh = py - 1;
    bi.w0 = phase * 4 + 0 * 2 + b;
    bi.w1 = phase * 4 + 1 * 2 + b;
  } else {
    const int v = box >> 1, a = box & 1;                      // input row = 2*oy - 1 + a + 2*bb, bb = 0,1
    bi.dw = v - 1;
    bi.dh = a - 1;
    bi.w0 = (a + 0) * 4 + v;
    bi.w1 = (a + 2) * 4 + v;
  }
  return bi;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
convud_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + (size_t)p.stages * p.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = blockIdx.x % p.tiles_n;
  const int sp0 = blockIdx.x / p.tiles_n, sp_step = gridDim.x / p.tiles_n;
  const int n0 = nt * p.BN;
  const int b_tap_bytes = p.BN * BK * 2;
  const int in_mul = (p.mode == MODE_DOWN) ? 2 : 1;
  __shared__ float s_bias[128];
  for (int i = threadIdx.x; i < p.BN; i += blockDim.x) s_bias[i] = p.bias ? p.bias[n0 + i] * p.bias_mul : 0.f;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bars->tmem_full[a], 1); mbar_init(&bars->tmem_empty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&map_x);
    prefetch_tmap(&map_w);
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
      for (int sp = sp0; sp < p.num_spatial; sp += sp_step) {
        int r = sp;
        const int wt = r % p.tiles_w; r /= p.tiles_w;
        const int ht = r % p.tiles_h;
        const int n = r / p.tiles_h;
        const int h0 = ht * TH * in_mul, w0 = wt * TW * in_mul;
        for (int ph = 0; ph < p.nphase; ++ph) {
          for (int kc = 0; kc < p.kchunks; ++kc) {
            for (int box = 0; box < p.nbox; ++box) {
              const BoxInfo bi = box_info(p.mode, ph, box);
              mbar_wait(&bars->empty[stage], phase ^ 1);
              uint8_t* sA = smem + (size_t)stage * p.stage_bytes;
              uint8_t* sB = sA + A_BYTES;
              mbar_expect_tx(&bars->full[stage], (uint32_t)(A_BYTES + 2 * b_tap_bytes));
              tma_load_4d(sA, &map_x, &bars->full[stage], kc * BK, w0 + bi.dw, h0 + bi.dh, n);
              tma_load_2d(sB, &map_w, &bars->full[stage], kc * BK, bi.w0 * p.Cout + n0);
              tma_load_2d(sB + b_tap_bytes, &map_w, &bars->full[stage], kc * BK, bi.w1 * p.Cout + n0);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform, one elected lane issues) =====================
    const uint32_t idesc = make_idesc_bf16(128, p.BN);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bstep = (uint32_t)b_tap_bytes >> 4;
    const int iters = p.kchunks * p.nbox;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int sp = sp0; sp < p.num_spatial; sp += sp_step) {
      for (int ph = 0; ph < p.nphase; ++ph) {
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.BN);
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_base + (uint32_t)(stage * p.stage_bytes);
          const uint64_t a0 = make_desc_sw128(sA);
          const uint64_t b0 = make_desc_sw128(sA + A_BYTES);
          const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32);
          const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
          const bool leader = elect_one_sync();
#pragma unroll
          for (int t = 0; t < 2; ++t) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint32_t al = a_lo + (uint32_t)((t * (TW * 128) + k * 32) >> 4);
              const uint32_t bl = b_lo + t * bstep + (uint32_t)((k * 32) >> 4);
              if (leader) umma_bf16_lohi(tmem_d, al, a_hi, bl, b_hi, idesc, (it | t | k) ? 1u : 0u);
            }
          }
          if (leader) umma_commit(&bars->empty[stage]);
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) umma_commit(&bars->tmem_full[acc]);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (TMEM lane quadrant = warp % 4) =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int pr = m / TW, pc = m % TW;
    const int out_mul = (p.mode == MODE_UP) ? 2 : 1;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int sp = sp0; sp < p.num_spatial; sp += sp_step) {
      int r = sp;
      const int wt = r % p.tiles_w; r /= p.tiles_w;
      const int ht = r % p.tiles_h;
      const int n = r / p.tiles_h;
      for (int ph = 0; ph < p.nphase; ++ph) {
        const int oy = (ht * TH + pr) * out_mul + (ph >> 1);
        const int ox = (wt * TW + pc) * out_mul + (ph & 1);
        mbar_wait(&bars->tmem_full[acc], acc_phase);
        tc_fence_after();
        __nv_bfloat16* dst = p.y + ((((size_t)n * p.Hout + oy) * p.Wout + ox) * p.Cout + n0);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
        auto emit16 = [&](const uint32_t (&v)[16], int c0) {
          uint32_t packed[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a = __uint_as_float(v[2 * j]) + s_bias[c0 + 2 * j];
            float b = __uint_as_float(v[2 * j + 1]) + s_bias[c0 + 2 * j + 1];
            if (p.act) { a = sg_lrelu(a); b = sg_lrelu(b); }
            __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
            packed[j] = *reinterpret_cast<uint32_t*>(&h2);
          }
          st_global_256(dst + c0, packed);
        };
        for (int c0 = 0; c0 < p.BN; c0 += 64) {
          uint32_t v0[16], v1[16], v2[16], v3[16];
          const int nld = (p.BN - c0) >= 64 ? 4 : (p.BN - c0) / 16;
          tmem_ld16_nowait(taddr + c0, v0);
          if (nld > 1) tmem_ld16_nowait(taddr + c0 + 16, v1);
          if (nld > 2) tmem_ld16_nowait(taddr + c0 + 32, v2);
          if (nld > 3) tmem_ld16_nowait(taddr + c0 + 48, v3);
          tmem_ld_wait();
          if (c0 + 64 >= p.BN) {                            // last chunk is in registers: release the TMEM buffer now
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
          }
          emit16(v0, c0);
          if (nld > 1) emit16(v1, c0 + 16);
          if (nld > 2) emit16(v2, c0 + 32);
          if (nld > 3) emit16(v3, c0 + 48);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}


// ---- CTA-pair (cta_group::2) form: M = 256 = the 8x16 tiles of two consecutive spatial positions, each CTA holding half of
// the weight rows (see sg_conv_pair_tc.cu for the protocol: TMA completions on the leader's mbarrier, multicast commits,
// remote tmem_empty arrivals).  1.6x the MMA issue rate of the single-CTA form at Cout = 64, 1.25x at 128.
struct __align__(8) PairBarriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
convud_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  PairBarriers* bars = reinterpret_cast<PairBarriers*>(smem + (size_t)p.stages * p.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1u;          // == %cluster_ctarank for a 1-D grid of 2-CTA clusters, and provably warp-uniform: with the special register the MMA loop lost its uniform-datapath form (an ELECT/R2UR waterfall per tcgen05.mma, ~110 cycles each)
  const bool leader_cta = rank == 0;
  const int cluster_id = blockIdx.x >> 1, clusters = gridDim.x >> 1;
  const int nt = cluster_id % p.tiles_n;
  const int q0 = cluster_id / p.tiles_n, q_step = clusters / p.tiles_n;
  const int Q = p.num_spatial / 2;
  const int n0 = nt * p.BN;
  const int HB = p.BN / 2;
  const int b_tap_bytes = HB * BK * 2;
  const int in_mul = (p.mode == MODE_DOWN) ? 2 : 1;
  __shared__ float s_bias[256];
  for (int i = threadIdx.x; i < p.BN; i += blockDim.x) s_bias[i] = p.bias ? p.bias[n0 + i] * p.bias_mul : 0.f;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bars->tmem_full[a], 1); mbar_init(&bars->tmem_empty[a], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&map_x);
    prefetch_tmap(&map_w);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      const int wrow0 = n0 + (int)rank * HB;
      const uint32_t stage_tx = (uint32_t)(2 * (A_BYTES + 2 * b_tap_bytes));
      int stage = 0;
      uint32_t phase = 0;
      for (int q = q0; q < Q; q += q_step) {
        int r = 2 * q + (int)rank;
        const int wt = r % p.tiles_w; r /= p.tiles_w;
        const int ht = r % p.tiles_h;
        const int n = r / p.tiles_h;
        const int h0 = ht * TH * in_mul, w0 = wt * TW * in_mul;
        for (int ph = 0; ph < p.nphase; ++ph) {
          for (int kc = 0; kc < p.kchunks; ++kc) {
            for (int box = 0; box < p.nbox; ++box) {
              const BoxInfo bi = box_info(p.mode, ph, box);
              mbar_wait(&bars->empty[stage], phase ^ 1);
              uint8_t* sA = smem + (size_t)stage * p.stage_bytes;
              uint8_t* sB = sA + A_BYTES;
              if (leader_cta) mbar_expect_tx(&bars->full[stage], stage_tx);
              tma2_load_4d(sA, &map_x, &bars->full[stage], kc * BK, w0 + bi.dw, h0 + bi.dh, n);
              tma2_load_2d(sB, &map_w, &bars->full[stage], kc * BK, bi.w0 * p.Cout + wrow0);
              tma2_load_2d(sB + b_tap_bytes, &map_w, &bars->full[stage], kc * BK, bi.w1 * p.Cout + wrow0);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader_cta) {
      const uint32_t idesc = make_idesc_bf16(256, p.BN);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t bstep = (uint32_t)b_tap_bytes >> 4;
      const int iters = p.kchunks * p.nbox;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int q = q0; q < Q; q += q_step) {
        for (int ph = 0; ph < p.nphase; ++ph) {
          mbar_wait_cluster(&bars->tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.BN);
          for (int it = 0; it < iters; ++it) {
            mbar_wait(&bars->full[stage], phase);
            tc_fence_after();
            const uint32_t sA = smem_base + (uint32_t)(stage * p.stage_bytes);
            const uint64_t a0 = make_desc_sw128(sA);
            const uint64_t b0 = make_desc_sw128(sA + A_BYTES);
            const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32);
            const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
            const bool leader = elect_one_sync();
#pragma unroll
            for (int t = 0; t < 2; ++t) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint32_t al = a_lo + (uint32_t)((t * (TW * 128) + k * 32) >> 4);
                const uint32_t bl = b_lo + t * bstep + (uint32_t)((k * 32) >> 4);
                if (leader) umma2_bf16_lohi(tmem_d, al, a_hi, bl, b_hi, idesc, (it | t | k) ? 1u : 0u);
              }
            }
            if (leader) umma2_commit(&bars->empty[stage]);
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          if (elect_one_sync()) umma2_commit(&bars->tmem_full[acc]);
          __syncwarp();
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    const int qd = warp & 3;
    const int m = qd * 32 + lane;
    const int pr = m / TW, pc = m % TW;
    const int out_mul = (p.mode == MODE_UP) ? 2 : 1;
    const uint32_t empty_remote = mapa_u32(smem_u32(&bars->tmem_empty[0]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int q = q0; q < Q; q += q_step) {
      int r = 2 * q + (int)rank;
      const int wt = r % p.tiles_w; r /= p.tiles_w;
      const int ht = r % p.tiles_h;
      const int n = r / p.tiles_h;
      for (int ph = 0; ph < p.nphase; ++ph) {
        const int oy = (ht * TH + pr) * out_mul + (ph >> 1);
        const int ox = (wt * TW + pc) * out_mul + (ph & 1);
        mbar_wait(&bars->tmem_full[acc], acc_phase);
        tc_fence_after();
        __nv_bfloat16* dst = p.y + ((((size_t)n * p.Hout + oy) * p.Wout + ox) * p.Cout + n0);
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(acc * p.BN);
        auto emit16 = [&](const uint32_t (&v)[16], int c0) {
          uint32_t packed[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a = __uint_as_float(v[2 * j]) + s_bias[c0 + 2 * j];
            float b = __uint_as_float(v[2 * j + 1]) + s_bias[c0 + 2 * j + 1];
            if (p.act) { a = sg_lrelu(a); b = sg_lrelu(b); }
            __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
            packed[j] = *reinterpret_cast<uint32_t*>(&h2);
          }
          st_global_256(dst + c0, packed);
        };
        for (int c0 = 0; c0 < p.BN; c0 += 64) {
          uint32_t v0[16], v1[16], v2[16], v3[16];
          tmem_ld16_nowait(taddr + c0, v0);
          tmem_ld16_nowait(taddr + c0 + 16, v1);
          tmem_ld16_nowait(taddr + c0 + 32, v2);
          tmem_ld16_nowait(taddr + c0 + 48, v3);
          tmem_ld_wait();
          if (c0 + 64 >= p.BN) {                            // last chunk is in registers: release the TMEM buffer now
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (leader_cta) mbar_arrive(&bars->tmem_empty[acc]);
              else mbar_arrive_remote(empty_remote + (uint32_t)(acc * 8));
            }
          }
          emit16(v0, c0);
          emit16(v1, c0 + 16);
          emit16(v2, c0 + 32);
          emit16(v3, c0 + 48);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

static int pick_bn_pair(int Cout) {
  if (Cout % 256 == 0) return 256;
  if (Cout == 128 || Cout == 64) return Cout;
  return 0;
}
static bool pair_ok(int N, int H, int W, int Cin, int Cout) {
  return pick_bn_pair(Cout) != 0 && Cin % BK == 0 && Cin >= BK && ((N * (H / TH) * (W / TW)) % 2 == 0);
}

static int pick_bn(int Cout) {
  if (Cout % 128 == 0) return 128;
  if (Cout <= 128 && Cout % 16 == 0) return Cout;
  return 0;
}

// 4-D NHWC bf16 map fetched with element strides 2 in W and H: box = box_w x box_h LOADED pixels.
static bool encode_nhwc_stride2(CUtensorMap* map, const void* ptr, int N, int H, int W, int C, int box_w, int box_h) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)(2 * box_w), (cuuint32_t)(2 * box_h), 1u};
  cuuint32_t es[4] = {1, 2, 2, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace ud

// mode 0 = UP (input [N,H,W,Cin] -> output [N,2H,2W,Cout]), mode 1 = DOWN (input [N,2H,2W,Cin] -> output [N,H,W,Cout]);
// H, W are the LOW-resolution dimensions in both modes.
SG_EXPORT int sg_convud_tc_supported(int mode, int N, int H, int W, int Cin, int Cout) {
  if (mode != 0 && mode != 1) return 0;
  if (N < 1 || H % ud::TH != 0 || W % ud::TW != 0) return 0;
  if (Cin < 16 || Cin % 8 != 0) return 0;      // 64-channel TMA boxes, zero-filled beyond Cin
  if (ud::pick_bn(Cout) == 0) return 0;
  return 1;
}

// E: [ntaps][Cout][Cin] bf16 with ntaps = 16 (UP: [phase][a*2+b]; DOWN: [u*4+v]).
SG_EXPORT int sg_convud_tc(int mode, const void* x, const void* E, void* y, const float* bias, float bias_mul, int act,
                           int N, int H, int W, int Cin, int Cout, void* stream) {
  using namespace ud;
  SG_REQUIRE(sg_convud_tc_supported(mode, N, H, W, Cin, Cout), "sg_convud_tc: shape not supported");
  if (g_sg_pair_updown >= 2 && sg_convud_halo_supported(mode, N, H, W, Cin, Cout))
    return sg_convud_halo(mode, x, E, y, bias, bias_mul, act, N, H, W, Cin, Cout, stream);
  EncodeTiledFn enc = get_encode();
  SG_REQUIRE(enc != nullptr, "sg_convud_tc: cuTensorMapEncodeTiled unavailable");
  const bool pair = g_sg_pair_updown && pair_ok(N, H, W, Cin, Cout);
  const int BN = pair ? pick_bn_pair(Cout) : pick_bn(Cout);
  CUtensorMap map_x, map_w;
  if (mode == MODE_UP) {
    SG_REQUIRE(encode_nhwc(&map_x, x, N, H, W, Cin, TW, TH + 1), "sg_convud_tc: cuTensorMapEncodeTiled(x) failed");
  } else {
    SG_REQUIRE(encode_nhwc_stride2(&map_x, x, N, 2 * H, 2 * W, Cin, TW, TH + 1),
               "sg_convud_tc: cuTensorMapEncodeTiled(x, stride 2) failed");
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)16 * Cout};
    cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(pair ? BN / 2 : BN)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(E), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SG_REQUIRE(r == CUDA_SUCCESS, "sg_convud_tc: cuTensorMapEncodeTiled(w) failed");
  }
  Params p;
  p.mode = mode; p.N = N; p.Hq = H; p.Wq = W; p.Cin = Cin; p.Cout = Cout; p.BN = BN;
  p.Hout = (mode == MODE_UP) ? 2 * H : H;
  p.Wout = (mode == MODE_UP) ? 2 * W : W;
  p.tiles_h = H / TH; p.tiles_w = W / TW; p.tiles_n = Cout / BN;
  p.num_spatial = N * p.tiles_h * p.tiles_w;
  p.nphase = (mode == MODE_UP) ? 4 : 1;
  p.nbox = (mode == MODE_UP) ? 2 : 8;
  p.kchunks = (Cin + BK - 1) / BK;
  p.stage_bytes = A_BYTES + 2 * (pair ? BN / 2 : BN) * BK * 2;
  int stages = (214 * 1024) / p.stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  SG_REQUIRE(stages >= 2, "sg_convud_tc: not enough shared memory for two stages");
  p.stages = stages;
  int cols = 2 * BN, tc_cols = 32;
  while (tc_cols < cols) tc_cols <<= 1;
  p.tmem_cols = tc_cols;
  p.bias = bias; p.bias_mul = bias_mul; p.act = act;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  size_t smem = (size_t)stages * p.stage_bytes + sizeof(Barriers) + 1024;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(convud_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(convud_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  });
  SG_REQUIRE(attr_err == cudaSuccess, "sg_convud_tc: cannot raise the dynamic shared memory limit");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (pair) {
    int clusters = ((sms / 2) / p.tiles_n) * p.tiles_n;
    const int Q = p.num_spatial / 2;
    if (clusters > Q * p.tiles_n) clusters = Q * p.tiles_n;
    SG_REQUIRE(clusters >= p.tiles_n, "sg_convud_tc: more N blocks than SM pairs");
    smem = (size_t)stages * p.stage_bytes + sizeof(PairBarriers) + 1024;
    convud_pair_kernel<<<2 * clusters, NUM_THREADS, smem, (cudaStream_t)stream>>>(map_x, map_w, p);
    return sg_check_launch("sg_convud_tc/pair");
  }
  int grid = (sms / p.tiles_n) * p.tiles_n;
  if (grid > p.num_spatial * p.tiles_n) grid = p.num_spatial * p.tiles_n;
  SG_REQUIRE(grid >= p.tiles_n, "sg_convud_tc: more N blocks than SMs");
  convud_tc_kernel<<<grid, NUM_THREADS, smem, (cudaStream_t)stream>>>(map_x, map_w, p);
  return sg_check_launch("sg_convud_tc");
}

// 2 (default): the halo-box CTA-pair kernel (csrc/sg_convud2_tc.cu) where the shape allows, else as 1; 1: the first
// CTA-pair kernel where the shape allows; 0: single-CTA kernel only; <0: query.  Returns the old value.
SG_EXPORT int sg_pair_updown(int enable) {
  int old = g_sg_pair_updown;
  if (enable >= 0) g_sg_pair_updown = enable > 2 ? 2 : enable;
  return old;
}
