// Second-generation up/down-sampling convolution (bf16, tcgen05 cta_group::2, fp32 TMEM): the tap-merged low-resolution
// form of csrc/sg_convud_tc.cu (same mathematics, same weight layout E[16][Cout][Cin], reference:
// models/CustomLayers.py:143-180) with the three things the 3x3 pair kernel taught (csrc/sg_conv_pair_tc.cu):
//
//  * ONE halo box serves several taps.  Tile = 16 x 8 low-resolution pixels (M = 128).
//      UP:   the 18 x 10 box of x around the tile; output phase (py, px), tap (a, b) reads x[i + a + py - 1, j + b + px - 1]
//            = descriptor start + ((a + py) * 10 + (b + px)) * 128, rows 1280 B apart: 4 taps per box.
//      DOWN: input row 2i - 1 + u: the four (row parity, column parity) sub-grids of the high-resolution input are fetched
//            as 17 x 9 boxes with TMA element strides 2; tap (u, v) lives in sub-grid (u & 1, v & 1) at offset
//            (u >> 1, v >> 1) = start + ((u >> 1) * 9 + (v >> 1)) * 128, rows 1152 B apart: 4 taps per box.
//    A ring slot = one box + the four weight taps that go with it (16 MMAs).  The first kernel fetched one box per
//    horizontal tap and vertical pair: 160 KB (UP) / 144 KB (DOWN) of activations per 64-channel chunk and tile instead
//    of 94 / 80 KB, and only 8 MMAs per barrier round trip.
//  * Several MMA-issuing warps (the MMA stream of one warp is issue-bound below N = 256) taking alternate accumulation
//    units (UP: one output phase of a tile; DOWN: a tile), 512 / BN accumulator buffers in TMEM.
//  * Two epilogue teams of four warps taking alternate units.
#include "sg_common.cuh"
#include "../../include/sgb200.h"
#include "sg_tc_common.cuh"

namespace ud2 {

using namespace tcu;

constexpr int TH = 16, TW = 8, BK = 64;
constexpr int UP_W = TW + 2, UP_H = TH + 2;        // 10 x 18 halo box
constexpr int DN_W = TW + 1, DN_H = TH + 1;        // 9 x 17 box of one parity sub-grid
constexpr int A_SLOT = 23 * 1024;                  // 18*10*128 = 23040 and 17*9*128 = 19584, slots 1024-byte aligned
constexpr int MAX_STAGES = 8, MAX_ACC = 8, MAX_MMA_WARPS = 2;
constexpr int NUM_THREADS = 320 + 32 * (MAX_MMA_WARPS - 1);   // TMA warp, MMA warp, 8 epilogue warps, further MMA warps
constexpr int MODE_UP = 0, MODE_DOWN = 1;

struct __align__(8) Barriers {
  uint64_t full[MAX_STAGES];      // leader's copy is used: expect_tx by the leader, complete_tx from both CTAs' TMA
  uint64_t empty[MAX_STAGES];     // each CTA's own copy, arrived by the leader's multicast tcgen05.commit
  uint64_t tmem_full[MAX_ACC];    // each CTA's own copy (multicast commit)
  uint64_t tmem_empty[MAX_ACC];   // leader copy: the 4 warps of the owning team of each CTA arrive (count 8)
  uint32_t tmem_base;
  uint32_t pad;
};

struct Params {
  int mode, N, Hq, Wq, Cin, Cout, BN, HB, Hout, Wout;
  int tiles_h, tiles_w, tiles_n, Q, pairs;        // Q = tile pairs per N block, pairs = clusters per N block
  int kchunks, stages, stage_bytes, tmem_cols, nacc, mma_warps;
  int units_per_tile, slots_per_unit;             // UP: 4 units of kchunks slots; DOWN: 1 unit of 4*kchunks slots
  const float* bias;
  float bias_mul;
  int act;
  __nv_bfloat16* y;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
convud_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + (size_t)p.stages * p.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1u;          // == %cluster_ctarank for a 1-D grid of 2-CTA clusters (warp-uniform)
  const bool leader_cta = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int nt = cluster_id % p.tiles_n, j = cluster_id / p.tiles_n;
  const int n0 = nt * p.BN;
  const int q_begin = (int)((long long)j * p.Q / p.pairs), q_end = (int)((long long)(j + 1) * p.Q / p.pairs);
  const int b_tap_bytes = p.HB * BK * 2;          // one tap of this CTA's half of the weights
  const bool up = p.mode == MODE_UP;

  __shared__ float s_bias[256];
  for (int i = threadIdx.x; i < p.BN; i += blockDim.x) s_bias[i] = p.bias ? p.bias[n0 + i] * p.bias_mul : 0.f;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int a = 0; a < MAX_ACC; ++a) { mbar_init(&bars->tmem_full[a], 1); mbar_init(&bars->tmem_empty[a], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&map_x);
    prefetch_tmap(&map_w);
  }
  if (warp == 1) {                                        // one warp in EACH CTA of the pair
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
  const int n_units = (q_end - q_begin) * p.units_per_tile;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      const int wrow0 = n0 + (int)rank * p.HB;
      const uint32_t stage_tx = (uint32_t)(2 * ((up ? UP_H * UP_W : DN_H * DN_W) * BK * 2 + 4 * b_tap_bytes));
      int stage = 0;
      uint32_t phase = 0;
      for (int q = q_begin; q < q_end; ++q) {
        int r = 2 * q + (int)rank;
        const int wt = r % p.tiles_w; r /= p.tiles_w;
        const int ht = r % p.tiles_h;
        const int n = r / p.tiles_h;
        const int h0 = ht * TH, w0 = wt * TW;
        for (int u = 0; u < p.units_per_tile; ++u) {
          for (int s = 0; s < p.slots_per_unit; ++s) {
            // UP: unit = output phase g, slot = chunk kc.  DOWN: slot = (kc, g), g = (row parity, column parity) of the sub-grid
            const int kc = up ? s : (s >> 2), g = up ? u : (s & 3);
            mbar_wait(&bars->empty[stage], phase ^ 1);
            uint8_t* sA = smem + (size_t)stage * p.stage_bytes;
            uint8_t* sB = sA + A_SLOT;
            if (leader_cta) mbar_expect_tx(&bars->full[stage], stage_tx);
            if (up) tma2_load_4d(sA, &map_x, &bars->full[stage], kc * BK, w0 - 1, h0 - 1, n);
            else tma2_load_4d(sA, &map_x, &bars->full[stage], kc * BK, 2 * w0 - 1 + (g & 1), 2 * h0 - 1 + (g >> 1), n);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              // weight tile of tap t of group g:  UP [phase][a*2+b];  DOWN [u*4+v], u = 2*(t>>1) + (g>>1), v = 2*(t&1) + (g&1)
              const int wt_idx = up ? (g * 4 + t) : ((2 * (t >> 1) + (g >> 1)) * 4 + 2 * (t & 1) + (g & 1));
              tma2_load_2d(sB + t * b_tap_bytes, &map_w, &bars->full[stage], kc * BK, wt_idx * p.Cout + wrow0);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || warp >= 10) {
    // ===================== MMA issuers (leader CTA only): alternate accumulation units =====================
    const int mw = warp == 1 ? 0 : warp - 9;
    if (leader_cta && mw < p.mma_warps) {
      const uint32_t idesc = make_idesc_bf16(256, p.BN);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t bstep = (uint32_t)b_tap_bytes >> 4;
      const uint32_t a_sbo = (uint32_t)((up ? UP_W : DN_W) * BK * 2);
      const int row_pitch = up ? UP_W : DN_W;
      int stage = 0;
      uint32_t phase = 0;
      auto skip_units = [&](int n) {
        stage += n * p.slots_per_unit;
        while (stage >= p.stages) { stage -= p.stages; phase ^= 1; }
      };
      skip_units(mw);
      for (int ui = mw; ui < n_units; ui += p.mma_warps) {
        const int acc = ui % p.nacc;
        const int g_unit = ui % p.units_per_tile;            // UP: output phase of this unit
        mbar_wait_cluster(&bars->tmem_empty[acc], ((uint32_t)(ui / p.nacc) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.BN);
        for (int s = 0; s < p.slots_per_unit; ++s) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_base + (uint32_t)(stage * p.stage_bytes);
          // UP: the box is shared by the four phases, the phase shifts the tap window by (py, px)
          const int oy = up ? (g_unit >> 1) : 0, ox = up ? (g_unit & 1) : 0;
          const uint64_t a0 = make_desc_sw128_sbo(sA + (uint32_t)((oy * row_pitch + ox) * 128), a_sbo);
          const uint64_t b0 = make_desc_sw128(sA + (uint32_t)A_SLOT);
          const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32);
          const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
          const uint32_t a_row = (uint32_t)((row_pitch * 128) >> 4);   // tap row t>>1 -> one box row down
          const bool leader = elect_one_sync();
#pragma unroll
          for (int t = 0; t < 4; ++t) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint32_t al = a_lo + (uint32_t)(t >> 1) * a_row + (uint32_t)(((t & 1) * 128 + k * 32) >> 4);
              const uint32_t bl = b_lo + t * bstep + (uint32_t)((k * 32) >> 4);
              if (leader) umma2_bf16_lohi(tmem_d, al, a_hi, bl, b_hi, idesc, (s | t | k) ? 1u : 0u);
            }
          }
          if (leader) umma2_commit(&bars->empty[stage]);
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) umma2_commit(&bars->tmem_full[acc]);
        __syncwarp();
        skip_units(p.mma_warps - 1);
      }
    }
  } else {
    // ===================== epilogue: two teams of four warps (alternate units), both CTAs =====================
    const int team = (warp - 2) >> 2;
    const int qd = warp & 3;                               // TMEM lane quadrant = warp % 4 (hardware rule)
    const int m = qd * 32 + lane;
    const int pr = m / TW, pc = m % TW;
    const uint32_t empty_remote = mapa_u32(smem_u32(&bars->tmem_empty[0]), 0);
    const bool has_bias = p.bias != nullptr;
    const int nchunks = p.BN / 64;
    for (int ui = team; ui < n_units; ui += 2) {
      const int q = q_begin + ui / p.units_per_tile, g = ui % p.units_per_tile;
      int r = 2 * q + (int)rank;
      const int wt = r % p.tiles_w; r /= p.tiles_w;
      const int ht = r % p.tiles_h;
      const int n = r / p.tiles_h;
      const int oy = up ? (ht * TH + pr) * 2 + (g >> 1) : ht * TH + pr;
      const int ox = up ? (wt * TW + pc) * 2 + (g & 1) : wt * TW + pc;
      __nv_bfloat16* dst = p.y + ((((size_t)n * p.Hout + oy) * p.Wout + ox) * p.Cout + n0);
      const int buf = ui % p.nacc;
      mbar_wait(&bars->tmem_full[buf], (uint32_t)(ui / p.nacc) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * p.BN);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c >= nchunks) break;
        const int c0 = c * 64;
        uint32_t v0[16], v1[16], v2[16], v3[16];
        tmem_ld16_nowait(taddr + c0, v0);
        tmem_ld16_nowait(taddr + c0 + 16, v1);
        tmem_ld16_nowait(taddr + c0 + 32, v2);
        tmem_ld16_nowait(taddr + c0 + 48, v3);
        tmem_ld_wait();
        if (c == nchunks - 1) {                             // the accumulator is in registers: hand the TMEM buffer back now
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader_cta) mbar_arrive(&bars->tmem_empty[buf]);
            else mbar_arrive_remote(empty_remote + (uint32_t)(buf * 8));
          }
        }
        auto emit16 = [&](const uint32_t (&v)[16], int o) {
          uint32_t packed[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            float a = __uint_as_float(v[2 * jj]), b = __uint_as_float(v[2 * jj + 1]);
            if (has_bias) { a += s_bias[c0 + o + 2 * jj]; b += s_bias[c0 + o + 2 * jj + 1]; }
            if (p.act) { a = fmaxf(a, SG_LRELU_SLOPE * a); b = fmaxf(b, SG_LRELU_SLOPE * b); }
            __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
            packed[jj] = *reinterpret_cast<uint32_t*>(&h2);
          }
          st_global_256(dst + c0 + o, packed);
        };
        emit16(v0, 0);
        emit16(v1, 16);
        emit16(v2, 32);
        emit16(v3, 48);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                     // the leader's MMAs read the peer's shared memory until here
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

static int pick_bn(int Cout) {
  if (Cout % 256 == 0) return 256;
  if (Cout == 128 || Cout == 64) return Cout;
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

}  // namespace ud2

// H, W are the LOW-resolution dimensions in both modes (as sg_convud_tc).
int sg_convud_halo_supported(int mode, int N, int H, int W, int Cin, int Cout) {
  using namespace ud2;
  if (mode != 0 && mode != 1) return 0;
  if (N < 1 || H % TH != 0 || W % TW != 0) return 0;
  if (Cin < BK || Cin % BK != 0 || pick_bn(Cout) == 0) return 0;
  if ((N * (H / TH) * (W / TW)) % 2 != 0) return 0;
  const int stage_bytes = A_SLOT + 4 * (pick_bn(Cout) / 2) * BK * 2;
  return (220 * 1024) / stage_bytes >= 2;
}

// same contract as sg_convud_tc (E: [16][Cout][Cin] bf16); called from there when the shape is supported
int sg_convud_halo(int mode, const void* x, const void* E, void* y, const float* bias, float bias_mul, int act, int N, int H,
                   int W, int Cin, int Cout, void* stream) {
  using namespace ud2;
  SG_REQUIRE(sg_convud_halo_supported(mode, N, H, W, Cin, Cout), "sg_convud_halo: shape not supported");
  EncodeTiledFn enc = get_encode();
  SG_REQUIRE(enc != nullptr, "sg_convud_halo: cuTensorMapEncodeTiled unavailable");
  const int BN = pick_bn(Cout);
  CUtensorMap map_x, map_w;
  if (mode == MODE_UP) {
    SG_REQUIRE(encode_nhwc(&map_x, x, N, H, W, Cin, UP_W, UP_H), "sg_convud_halo: cuTensorMapEncodeTiled(x) failed");
  } else {
    SG_REQUIRE(encode_nhwc_stride2(&map_x, x, N, 2 * H, 2 * W, Cin, DN_W, DN_H),
               "sg_convud_halo: cuTensorMapEncodeTiled(x, stride 2) failed");
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)16 * Cout};
    cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(BN / 2)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(E), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SG_REQUIRE(r == CUDA_SUCCESS, "sg_convud_halo: cuTensorMapEncodeTiled(w) failed");
  }
  Params p;
  p.mode = mode; p.N = N; p.Hq = H; p.Wq = W; p.Cin = Cin; p.Cout = Cout; p.BN = BN; p.HB = BN / 2;
  p.Hout = (mode == MODE_UP) ? 2 * H : H;
  p.Wout = (mode == MODE_UP) ? 2 * W : W;
  p.tiles_h = H / TH; p.tiles_w = W / TW; p.tiles_n = Cout / BN;
  p.Q = N * p.tiles_h * p.tiles_w / 2;
  p.kchunks = Cin / BK;
  p.units_per_tile = (mode == MODE_UP) ? 4 : 1;
  p.slots_per_unit = (mode == MODE_UP) ? p.kchunks : 4 * p.kchunks;
  p.stage_bytes = A_SLOT + 4 * p.HB * BK * 2;
  int stages = (220 * 1024) / p.stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  p.stages = stages;
  int nacc = 512 / BN;
  if (nacc > MAX_ACC) nacc = MAX_ACC;
  p.nacc = nacc;
  int cols = 32;
  while (cols < nacc * BN) cols <<= 1;
  p.tmem_cols = cols;
  {
    static const int mmaw = getenv("SG_PAIR_MMAW") ? atoi(getenv("SG_PAIR_MMAW")) : MAX_MMA_WARPS;   // bench only
    p.mma_warps = mmaw < 1 ? 1 : (mmaw > MAX_MMA_WARPS ? MAX_MMA_WARPS : mmaw);
    // an issuer that starts a unit ahead must never wait on the second fill of a ring slot (mbarrier parity waits cannot
    // tell "two phases ahead" from "done"): the ring has to hold the units in flight
    if (p.stages < p.mma_warps * p.slots_per_unit) p.mma_warps = 1;
  }
  p.bias = bias; p.bias_mul = bias_mul; p.act = act;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int pairs = (sms / 2) / p.tiles_n;
  if (pairs > p.Q) pairs = p.Q;
  SG_REQUIRE(pairs >= 1, "sg_convud_halo: more N blocks than SM pairs");
  p.pairs = pairs;
  const size_t smem = (size_t)p.stages * p.stage_bytes + sizeof(Barriers) + 1024;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(convud_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  });
  SG_REQUIRE(attr_err == cudaSuccess, "sg_convud_halo: cannot raise the dynamic shared memory limit");
  convud_halo_kernel<<<2 * pairs * p.tiles_n, NUM_THREADS, smem, (cudaStream_t)stream>>>(map_x, map_w, p);
  return sg_check_launch("sg_convud_halo");
}
