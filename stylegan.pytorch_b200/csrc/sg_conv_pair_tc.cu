// CTA-pair (tcgen05 cta_group::2) implicit-GEMM 3x3 stride-1 convolution for sm_100a with the synthesis-layer epilogue
// fused in:
//
//   z[n,h,w,co] = act( sum_{kh,kw,ci} x[n,h+kh-1,w+kw-1,ci] * E[kh*3+kw][co][ci] + bias[co]*bias_mul
//                      + noise_w[co] * noise[n,h,w] )                                     (bf16, written once)
//   part[n][slot][0][co] = sum over the CTA's pixels of image n of z,   part[n][slot][1][co] = same of z^2
//
// reference call sites: EqualizedConv2d.forward -> F.conv2d (models/CustomLayers.py:170-171) followed by the first
// half of LayerEpilogue.forward (NoiseLayer :191-200, LeakyReLU, and the statistics nn.InstanceNorm2d :233 needs);
// with noise/part null it is the plain fused-bias/LeakyReLU convolution and its dgrad (discriminator, backward).
//
// Why a CTA pair: a 128xNx16 cta_group::1 MMA never issues faster than one per ~61-68 cycles on this part whatever N is
// (profiles/r01_mma_rate.txt), which caps the 64-output-channel layers at 0.58 of the measured bf16 peak; the
// cta_group::2 form (M = 256: 128 pixels from each CTA, each CTA holding N/2 rows of the weights) measured 43 cycles at
// N = 64 and 64 at N = 128 (profiles/r02_mma_rate2.txt: 1573 / 2046 TFLOP/s chip-wide with operands resident).
//
// Structure per CTA (both CTAs of the cluster): warp 0 TMA producer (own 18x10 activation halo box per 64-channel chunk
// + own half of the weights, completion signalled on the LEADER's mbarrier via the cta_group::2 TMA form), warps 1 and 10
// MMA issuers (leader CTA only; alternate tiles -- one warp cannot issue tcgen05.mma fast enough, see below), warps 2-9
// two epilogue teams for the CTA's own 128-pixel (16 x 8) tiles.  A pair walks a CONTIGUOUS range of tile pairs, so the per-image
// statistics are accumulated in registers across tiles and flushed once per (CTA, image): deterministic two-stage
// reduction, no atomics (sg_epi_finalize_parts sums the slots in fixed order).
#include "sg_common.cuh"
#include "../../include/sgb200.h"
#include "sg_tc_common.cuh"
#include <stdlib.h>

namespace tc2 {

using namespace tcu;

// Tile = 16 rows x 8 pixels (M = 128: 8-pixel tile row r = rows 8r..8r+7 of the MMA).  ONE halo box of 18 x 10 pixels per
// 64-channel chunk serves all nine taps: TMA writes it as 180 dense 128-byte rows (128B swizzle), tap (kh, kw) is the
// descriptor start address + (kh*10 + kw)*128 with a stride of 1280 B between the 8-row groups.  The tensor core applies
// the swizzle to the absolute shared-memory address, so start addresses and group strides that are not multiples of the
// 1024-byte swizzle atom read correctly with base_offset = 0 (tools/umma_shift_probe.cu, profiles/r02_umma_shift_probe.txt).
// The first version loaded three kw-shifted 10 x 16 boxes (60 KB per chunk instead of 23 KB) to keep every descriptor
// 1024-byte aligned.
constexpr int TH = 16, TW = 8, BK = 64;
constexpr int HALO_W = TW + 2, HALO_H = TH + 2;
constexpr int A_BOX_BYTES = HALO_H * HALO_W * BK * 2;   // 23040
constexpr int A_BYTES = (A_BOX_BYTES + 1023) / 1024 * 1024;   // 23552: ring slots stay 1024-byte aligned
constexpr int A_SBO = HALO_W * BK * 2;             // 1280 B between consecutive tile rows in the box
constexpr int MAX_MMA_WARPS = 2;
constexpr int NUM_THREADS = 320 + 32 * (MAX_MMA_WARPS - 1);   // TMA warp, MMA warp, 8 epilogue warps, further MMA warps
constexpr int MAX_STAGES = 8;                      // activation ring slots (one 64-channel chunk of one tile each)
constexpr int MAX_BSTAGES = 4;                     // streamed-weight ring slots (three vertical taps of one (chunk, kw) each)

// Accumulator buffers in TMEM: nacc = 512 / BN (8 at BN = 64).  With two, the MMAs of tile q+2 wait for the whole
// full -> epilogue -> empty hand-over of tile q (a chain of mbarrier hops across the CTA pair, ~3000 cycles even with an
// empty epilogue) on top of the ~1550 cycles of tile q+1's MMAs: (L + T)/2 per tile.  More buffers let the MMA warp run
// ahead of the epilogue teams, which still take every other tile each.
constexpr int MAX_ACC = 8;
struct __align__(8) Barriers {
  uint64_t full[MAX_STAGES];      // leader's copy is used: expect_tx by the leader, complete_tx from both CTAs' TMA
  uint64_t empty[MAX_STAGES];     // each CTA's own copy, arrived by the leader's multicast tcgen05.commit
  uint64_t bfull[MAX_BSTAGES];    // the same pair for the streamed-weight ring (unused when the weights are resident)
  uint64_t bempty[MAX_BSTAGES];
  uint64_t tmem_full[MAX_ACC];    // each CTA's own copy (multicast commit)
  uint64_t tmem_empty[MAX_ACC];   // leader copy: the 4 warps of the owning team of each CTA arrive (count 8)
  uint64_t wfull;                 // leader's copy: resident weights of both CTAs have landed
  uint32_t tmem_base;
  uint32_t pad;
};

struct Params {
  int N, H, W, Cin, Cout, BN, HB;                 // BN = MMA N (both halves), HB = BN/2 = weight rows per CTA
  int tiles_h, tiles_w, tiles_n, num_spatial, tiles_per_image;
  int Q, pairs;                                   // tile pairs per N block, clusters per N block
  int kchunks, stages, stage_bytes, bstages, tmem_cols, nacc, mma_warps;   // stage_bytes = A_BYTES; bstages = 0 if resident
  int resident, w_bytes;
  const float* bias;
  float bias_mul;
  int act;
  const float* noise;                             // [N,H,W] fp32 or null
  const float* noise_w;                           // [Cout]
  int dbg;                                        // bench only (SG_PAIR_DBG): 1 = no global stores, 2 = no TMEM loads, 4 = empty epilogue, 8 = no statistics pass, 16 = no bias / noise terms, 32 = no activation TMA loads after the first tile
  float* part;                                    // [N][slots][2][Cout] fp32 or null; slots = 2*pairs
  __nv_bfloat16* y;
};

// statistics layout: after the per-warp column sums lane l holds channels 2l and 2l+1 of the 64-channel chunk
__device__ __forceinline__ int colsum_channel(int lane, int j) { return 2 * lane + j; }

template <bool STATS, bool N64>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_y, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem + (size_t)p.stages * p.stage_bytes;
  Barriers* bars = reinterpret_cast<Barriers*>(smem_w + p.w_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1u;          // == %cluster_ctarank for a 1-D grid of 2-CTA clusters, and provably warp-uniform: with the special register the MMA loop lost its uniform-datapath form (an ELECT/R2UR waterfall per tcgen05.mma, ~110 cycles each)
  const bool leader_cta = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int nt = cluster_id % p.tiles_n;                 // this pair's N block
  const int j = cluster_id / p.tiles_n;                  // pair index inside the N block
  const int n0 = nt * p.BN;
  // contiguous range of tile pairs
  const int q_begin = (int)((long long)j * p.Q / p.pairs), q_end = (int)((long long)(j + 1) * p.Q / p.pairs);

  __shared__ float s_bias[256], s_nw[256];
  __shared__ float s_red[8][2][64];                      // cross-warp combine of the statistics (one 64-ch chunk)
  // per-warp staging of one 32-pixel x 64-channel bf16 block (4 KB, 16-byte chunks XOR-swizzled by the pixel row so that
  // both the row-wise 16-byte stores and the column-wise 4-byte loads are bank-conflict free).  Column sums through
  // shared memory cost ~64 wavefronts per warp and chunk; a shuffle butterfly (tried first: 248 SHFL per warp and
  // chunk) made the epilogue 3x slower than the MMAs it has to hide behind.
  __shared__ __align__(1024) uint8_t s_stage[STATS ? 8 * 4096 : 16];   // TMA-store source: 128B-swizzle atoms
  for (int i = threadIdx.x; i < p.BN; i += blockDim.x) {
    s_bias[i] = p.bias ? p.bias[n0 + i] * p.bias_mul : 0.f;
    s_nw[i] = (p.noise && p.noise_w) ? p.noise_w[n0 + i] : 0.f;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int s = 0; s < p.bstages; ++s) { mbar_init(&bars->bfull[s], 1); mbar_init(&bars->bempty[s], 1); }
    for (int a = 0; a < MAX_ACC; ++a) { mbar_init(&bars->tmem_full[a], 1); mbar_init(&bars->tmem_empty[a], 8); }    // one team of 4 epilogue warps x 2 CTAs
    mbar_init(&bars->wfull, 1);
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
  cluster_sync_all();                                     // both CTAs' barriers initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int b_tap_bytes = p.HB * BK * 2;                  // one tap of this CTA's half of the weights

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      const int wrow0 = n0 + (int)rank * p.HB;
      if (p.resident) {
        if (leader_cta) mbar_expect_tx(&bars->wfull, (uint32_t)(2 * p.w_bytes));
        for (int kc = 0; kc < p.kchunks; ++kc)
          for (int t = 0; t < 9; ++t)
            tma2_load_2d(smem_w + (size_t)(kc * 9 + t) * b_tap_bytes, &map_w, &bars->wfull, kc * BK, t * p.Cout + wrow0);
      }
      int stage = 0, bstage = 0;
      uint32_t phase = 0, bphase = 0;
      for (int q = q_begin; q < q_end; ++q) {
        int r = 2 * q + (int)rank;
        int wt = r % p.tiles_w; r /= p.tiles_w;
        int ht = r % p.tiles_h;
        int n = r / p.tiles_h;
        const int h0 = ht * TH, w0 = wt * TW;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          if ((p.dbg & 32) && q > q_begin) {                  // probe: no activation loads after the first tile (stale operands)
            if (leader_cta) mbar_arrive(&bars->full[stage]);
          } else {
            if (leader_cta) mbar_expect_tx(&bars->full[stage], (uint32_t)(2 * A_BOX_BYTES));
            tma2_load_4d(smem + (size_t)stage * A_BYTES, &map_x, &bars->full[stage], kc * BK, w0 - 1, h0 - 1, n);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          if (!p.resident) {
            for (int kw = 0; kw < 3; ++kw) {
              mbar_wait(&bars->bempty[bstage], bphase ^ 1);
              uint8_t* sB = smem_w + (size_t)bstage * (3 * b_tap_bytes);
              if (leader_cta) mbar_expect_tx(&bars->bfull[bstage], (uint32_t)(2 * 3 * b_tap_bytes));
              for (int kh = 0; kh < 3; ++kh)
                tma2_load_2d(sB + kh * b_tap_bytes, &map_w, &bars->bfull[bstage], kc * BK, (kh * 3 + kw) * p.Cout + wrow0);
              if (++bstage == p.bstages) { bstage = 0; bphase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp >= 10) {
    // ===================== MMA issuers (leader CTA only) =====================
    // The MMA stream is ISSUE-bound, not tensor-bound: one warp needs ~65 cycles per tcgen05.mma (descriptor arithmetic
    // on the uniform datapath, R2UR moves, barrier waits) while the pair executes a 256x64x16 MMA in 43; issuing every MMA
    // twice cost +7 % (profiles/r02_summary.md).  So p.mma_warps warps (1 and 10, 11, ..., different scheduler quadrants)
    // issue the MMAs of alternate tiles, each into its own accumulator buffers; the shared-memory stages are consumed in
    // ring order by whichever warp owns the tile, and every commit tracks the issuing thread's own MMAs.
    const int mw = warp == 1 ? 0 : warp - 9;                // issuer index
    if (leader_cta && mw < p.mma_warps) {
      const uint32_t idesc = make_idesc_bf16(256, p.BN);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t sW = smem_u32(smem_w);
      const int per_tile = p.kchunks;                        // activation ring slots per tile (weight ring: 3 per chunk)
      int stage = 0, bstage = 0;
      uint32_t phase = 0, bphase = 0;
      auto skip_tiles = [&](int n) {                         // advance the ring positions past n tiles of other issuers
        stage += n * per_tile;
        while (stage >= p.stages) { stage -= p.stages; phase ^= 1; }
        if (!p.resident) {
          bstage += n * per_tile * 3;
          while (bstage >= p.bstages) { bstage -= p.bstages; bphase ^= 1; }
        }
      };
      skip_tiles(mw);
      if (p.resident) { mbar_wait(&bars->wfull, 0); tc_fence_after(); }
      const uint32_t bstep = (p.resident ? (uint32_t)(3 * b_tap_bytes) : (uint32_t)b_tap_bytes) >> 4;   // kh -> kh + 1
      for (int ti = mw; ti < q_end - q_begin; ti += p.mma_warps) {
        const int acc = ti % p.nacc;
        mbar_wait_cluster(&bars->tmem_empty[acc], ((uint32_t)(ti / p.nacc) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.BN);
        for (int kc = 0; kc < per_tile; ++kc) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint64_t a0 = make_desc_sw128_sbo(smem_base + (uint32_t)(stage * A_BYTES), (uint32_t)A_SBO);
          const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32);
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            uint32_t sB;
            if (p.resident) sB = sW + (uint32_t)((kc * 9 + kw) * b_tap_bytes);
            else {
              mbar_wait(&bars->bfull[bstage], bphase);
              tc_fence_after();
              sB = sW + (uint32_t)(bstage * 3 * b_tap_bytes);
            }
            const uint64_t b0 = make_desc_sw128(sB);
            const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
            const bool leader = elect_one_sync();
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint32_t al = a_lo + (uint32_t)(((kh * HALO_W + kw) * 128 + k * 32) >> 4);
                const uint32_t bl = b_lo + kh * bstep + (uint32_t)((k * 32) >> 4);
                if (leader) umma2_bf16_lohi(tmem_d, al, a_hi, bl, b_hi, idesc, (kc | kw | kh | k) ? 1u : 0u);
              }
            }
            if (!p.resident) {
              if (leader) umma2_commit(&bars->bempty[bstage]);
              if (++bstage == p.bstages) { bstage = 0; bphase ^= 1; }
            }
            __syncwarp();
          }
          if (elect_one_sync()) umma2_commit(&bars->empty[stage]);
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) umma2_commit(&bars->tmem_full[acc]);
        __syncwarp();
        skip_tiles(p.mma_warps - 1);
      }
    }
  } else {
    // ===================== epilogue: two teams of four warps, both CTAs =====================
    // TMEM lane quadrant = warp % 4 (hardware rule).  Team t = (warp - 2) / 4 owns accumulator buffer t, i.e. every
    // other tile of the CTA: the epilogue of a tile is a chain of fixed latencies (tmem_full hand-over, TMEM load,
    // shared-memory round trip of the statistics, global stores) of ~3600 cycles, longer than the ~2500 the MMAs of a
    // 64->64 tile pair take, and it does not shrink when the tile is split between more warps (tried: column halves).
    // Two teams overlap the chains of consecutive tiles instead.
    const int team = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int m = qd * 32 + lane;
    const int pr = m / TW, pc = m % TW;
    const uint32_t empty_remote = mapa_u32(smem_u32(&bars->tmem_empty[0]), 0);
    const int slot = 2 * j + (int)rank;
    const int slots = 2 * p.pairs;
    if constexpr (STATS && N64) {
      // ---- BN = 64 with statistics: everything per-channel lives in registers ---------------------------------------
      // The accumulator is read with the 16x256b TMEM shape: a thread gets, for 4 of the warp's 32 pixels, the SAME 16
      // channels (8k + 2*(lane&3) + {0,1}, k = 0..7).  So bias and noise weight of its channels are 32 registers loaded
      // once per kernel, and the InstanceNorm sums of its (pixels x channels) are 32 registers carried across ALL tiles
      // of an image: no per-tile column-sum pass and no per-value shared-memory read at all (the 32x32b form -- one
      // pixel, 64 channels per thread -- spent 2/3 of the epilogue's shared-memory wavefronts on broadcast reads of
      // bias / noise weight and on the column sums; profiles/r02_summary.md).  The cross-lane reduction (8 pixel groups)
      // happens once per image in flush64.
      const int rg = lane >> 2, cq = lane & 3;
      float bia[16], nwv[16], ssum[16], ssq[16];
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = n0 + 8 * k + 2 * cq + e;
          bia[2 * k + e] = (p.bias && !(p.dbg & 16)) ? p.bias[col] * p.bias_mul : 0.f;
          nwv[2 * k + e] = (p.noise && p.noise_w && !(p.dbg & 16)) ? p.noise_w[col] : 0.f;
          ssum[2 * k + e] = 0.f; ssq[2 * k + e] = 0.f;
        }
      int cur_n = -1;
      uint8_t* const sblk = s_stage + (team * 4 + qd) * 4096;   // this warp's 32 px x 64 ch bf16 block (TMA-store source)

      auto flush64 = [&](int n) {
        // 8 pixel groups (lane bits 2..4) -> lanes 0..3, fixed butterfly order; then the 8 warps through shared memory
#pragma unroll
        for (int i = 0; i < 16; ++i) {
#pragma unroll
          for (int d = 4; d < 32; d <<= 1) {
            ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], d);
            ssq[i] += __shfl_xor_sync(0xffffffffu, ssq[i], d);
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (rg == 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              s_red[team * 4 + qd][0][8 * k + 2 * cq + e] = ssum[2 * k + e];
              s_red[team * 4 + qd][1][8 * k + 2 * cq + e] = ssq[2 * k + e];
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int t = (warp - 2) * 32 + lane;
        if (t < 128) {
          const int which = t >> 6, ch = t & 63;
          float sum = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) sum += s_red[w8][which][ch];
          p.part[(((size_t)n * slots + slot) * 2 + which) * p.Cout + n0 + ch] = sum;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
      };

      // noise of this thread's 4 pixels (i = 16h + 8jp + rg of the warp's 32): tile rows 4qd + 2h + jp, column rg
      auto noise_of = [&](int q, float (&nz)[2][2]) {
        int r = 2 * q + (int)rank;
        const int wt = r % p.tiles_w; r /= p.tiles_w;
        const int ht = r % p.tiles_h;
        const int n = r / p.tiles_h;
        const float* src = p.noise + ((size_t)n * p.H + (ht * TH + 4 * qd)) * p.W + wt * TW + rg;
        nz[0][0] = src[0]; nz[0][1] = src[p.W]; nz[1][0] = src[2 * p.W]; nz[1][1] = src[3 * p.W];
      };
      const bool use_noise = p.noise != nullptr && p.noise_w != nullptr && !(p.dbg & 16);
      float nz_next[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
      if (use_noise && q_begin + team < q_end) noise_of(q_begin + team, nz_next);
      for (int q = q_begin; q < q_end; ++q) {
        const int n = (2 * q + (int)rank) / p.tiles_per_image;
        if (n != cur_n && cur_n >= 0) flush64(cur_n);          // both teams see every image boundary, in the same order
        cur_n = n;
        if (((q - q_begin) & 1) != team) continue;            // the other team's tile (accumulator buffer = team)
        const int rt = (2 * q + (int)rank) % p.tiles_per_image;
        const int tile_w0 = (rt % p.tiles_w) * TW, tile_h0 = (rt / p.tiles_w) * TH;
        float nz[2][2];
#pragma unroll
        for (int h = 0; h < 2; ++h) { nz[h][0] = nz_next[h][0]; nz[h][1] = nz_next[h][1]; }
        if (use_noise && q + 2 < q_end) noise_of(q + 2, nz_next);
        const int ti = q - q_begin, buf = ti % p.nacc;        // accumulator buffer of this tile; phase = its use count
        mbar_wait(&bars->tmem_full[buf], (uint32_t)(ti / p.nacc) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * 64);
        uint32_t v[2][32];
        if (!(p.dbg & 6)) {
          tmem_ld_16x256b_x8_nowait(taddr, v[0]);
          tmem_ld_16x256b_x8_nowait(taddr + (16u << 16), v[1]);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) { v[0][i] = 0u; v[1][i] = 0u; }
        }
        // the accumulator is in registers: hand the TMEM buffer back now
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader_cta) mbar_arrive(&bars->tmem_empty[buf]);
          else mbar_arrive_remote(empty_remote + (uint32_t)(buf * 8));
        }
        if (!(p.dbg & 4)) {
          // the previous TMA store out of this warp's staging block must have finished READING it
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {                  // pixel i = 16h + 8jp + rg of the warp's 32
              const float z = nz[h][jp];
              uint8_t* srow = sblk + (16 * h + 8 * jp + rg) * 128 + cq * 4;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                float a = __uint_as_float(v[h][4 * k + 2 * jp]) + bia[2 * k];
                float b = __uint_as_float(v[h][4 * k + 2 * jp + 1]) + bia[2 * k + 1];
                a = fmaf(nwv[2 * k], z, a);
                b = fmaf(nwv[2 * k + 1], z, b);
                if (p.act) { a = fmaxf(a, SG_LRELU_SLOPE * a); b = fmaxf(b, SG_LRELU_SLOPE * b); }
                __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                const uint32_t w = *reinterpret_cast<uint32_t*>(&h2);
                // 128B-swizzle atom: 16-byte chunk k of pixel row i sits at chunk position k ^ (i & 7) = k ^ rg
                *reinterpret_cast<uint32_t*>(srow + ((k ^ rg) << 4)) = w;
                if (!(p.dbg & 8)) {                           // the ROUNDED values, as the consumer will read them
                  const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
                  ssum[2 * k] += lo; ssum[2 * k + 1] += hi;
                  ssq[2 * k] = fmaf(lo, lo, ssq[2 * k]); ssq[2 * k + 1] = fmaf(hi, hi, ssq[2 * k + 1]);
                }
              }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0 && !(p.dbg & 1)) {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&map_y)), "r"(smem_u32(sblk)), "r"(n0), "r"(tile_w0),
                           "r"(tile_h0 + 4 * qd), "r"(n)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      if (cur_n >= 0) flush64(cur_n);
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
    } else {
    // running statistics of the current image (this team's tiles): lane l holds channels 2l, 2l+1 of each 64-ch chunk
    float st_s[4][2], st_q[4][2];
#pragma unroll
    for (int c = 0; c < 4; ++c) { st_s[c][0] = st_s[c][1] = st_q[c][0] = st_q[c][1] = 0.f; }
    int cur_n = -1;
    const int nchunks = p.BN / 64;
    uint8_t* const sblk = s_stage + (STATS ? (team * 4 + qd) * 4096 : 0);   // this warp's 32 px x 64 ch bf16 block

    auto flush = [&](int n) {
      // combine both teams' four lane quadrants through shared memory in a fixed order; write this CTA's slot of image n
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c >= nchunks) break;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        s_red[team * 4 + qd][0][2 * lane] = st_s[c][0]; s_red[team * 4 + qd][0][2 * lane + 1] = st_s[c][1];
        s_red[team * 4 + qd][1][2 * lane] = st_q[c][0]; s_red[team * 4 + qd][1][2 * lane + 1] = st_q[c][1];
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int t = (warp - 2) * 32 + lane;              // 256 threads, the first 128: 2 quantities x 64 channels
        if (t < 128) {
          const int which = t >> 6, ch = t & 63;
          float sum = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) sum += s_red[w8][which][ch];
          p.part[(((size_t)n * slots + slot) * 2 + which) * p.Cout + n0 + c * 64 + ch] = sum;
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) { st_s[c][0] = st_s[c][1] = st_q[c][0] = st_q[c][1] = 0.f; }
    };

    const bool has_bias = p.bias != nullptr && !(p.dbg & 16);
    const bool has_noise = p.noise != nullptr && p.noise_w != nullptr && !(p.dbg & 16);
    auto pixel_of = [&](int q) -> size_t {                 // this thread's pixel in the CTA's tile of pair q
      int r = 2 * q + (int)rank;
      const int wt = r % p.tiles_w; r /= p.tiles_w;
      const int ht = r % p.tiles_h;
      const int n = r / p.tiles_h;
      return ((size_t)n * p.H + (ht * TH + pr)) * p.W + (wt * TW + pc);
    };
    // the noise value of the team's NEXT tile is requested while the current one is processed
    const int q_first = q_begin + team;
    float nz_next = (p.noise && q_first < q_end) ? p.noise[pixel_of(q_first)] : 0.f;
    for (int q = q_begin; q < q_end; ++q) {
      const int n = (2 * q + (int)rank) / p.tiles_per_image;
      if (STATS) {                                          // both teams see every image boundary, in the same order
        if (n != cur_n && cur_n >= 0) flush(cur_n);
        cur_n = n;
      }
      if (((q - q_begin) & 1) != team) continue;            // the other team's tile (accumulator buffer = team)
      const size_t pix = pixel_of(q);
      const int rt = (2 * q + (int)rank) % p.tiles_per_image;
      const int tile_w0 = (rt % p.tiles_w) * TW, tile_h0 = (rt / p.tiles_w) * TH;
      const float nz = nz_next;
      if (p.noise && q + 2 < q_end) nz_next = p.noise[pixel_of(q + 2)];
      const int ti = q - q_begin, buf = ti % p.nacc;          // accumulator buffer of this tile; phase = its use count
      mbar_wait(&bars->tmem_full[buf], (uint32_t)(ti / p.nacc) & 1u);
      tc_fence_after();
      __nv_bfloat16* dst = p.y + pix * p.Cout + n0;
      const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * p.BN);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c >= nchunks || (p.dbg & 4)) break;
        const int c0 = c * 64;
        uint32_t v0[16], v1[16], v2[16], v3[16];
        if (!(p.dbg & 2)) {
          tmem_ld16_nowait(taddr + c0, v0);
          tmem_ld16_nowait(taddr + c0 + 16, v1);
          tmem_ld16_nowait(taddr + c0 + 32, v2);
          tmem_ld16_nowait(taddr + c0 + 48, v3);
          tmem_ld_wait();
        }
        if (c == nchunks - 1) {
          // the accumulator is in registers: hand the TMEM buffer back NOW, so that the MMAs of the tile after next do
          // not wait for this tile's arithmetic, stores and statistics
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader_cta) mbar_arrive(&bars->tmem_empty[buf]);
            else mbar_arrive_remote(empty_remote + (uint32_t)(buf * 8));
          }
        }
        uint8_t* srow = sblk + lane * 128;
        auto emit16 = [&](const uint32_t (&v)[16], int o) {
          uint32_t packed[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            float a = __uint_as_float(v[2 * jj]), b = __uint_as_float(v[2 * jj + 1]);
            if (has_bias) {                                 // (data-gradient launches carry no bias: no shared-memory reads)
              a += s_bias[c0 + o + 2 * jj];
              b += s_bias[c0 + o + 2 * jj + 1];
            }
            if (has_noise) {
              a = fmaf(s_nw[c0 + o + 2 * jj], nz, a);
              b = fmaf(s_nw[c0 + o + 2 * jj + 1], nz, b);
            }
            if (p.act) { a = fmaxf(a, SG_LRELU_SLOPE * a); b = fmaxf(b, SG_LRELU_SLOPE * b); }
            __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
            packed[jj] = *reinterpret_cast<uint32_t*>(&h2);
          }
          if (!STATS && !(p.dbg & 1)) st_global_256(dst + c0 + o, packed);
          if (STATS) {                                      // the ROUNDED values, as the consumer will read them
            const int ch = o >> 3;                          // logical 16-byte chunk of this pixel row (8 channels each)
            *reinterpret_cast<uint4*>(srow + (((ch + 0) ^ (lane & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            *reinterpret_cast<uint4*>(srow + (((ch + 1) ^ (lane & 7)) << 4)) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
          }
        };
        if (STATS) {
          // the previous TMA store out of this warp's staging block must have finished READING it
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
        emit16(v0, 0);
        emit16(v1, 16);
        emit16(v2, 32);
        emit16(v3, 48);
        if (STATS) {
          // The staged block IS the output: 32 pixels (four rows of the 16x8 tile) x 64 channels in the 128B-swizzle
          // layout.  One TMA store per warp and chunk replaces 4 x st.global.v8 per thread, whose 32-byte pieces landed in
          // 32 different 128-byte lines per instruction (512 LSU cycles per tile: the epilogue's largest single cost).
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0 && !(p.dbg & 1)) {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&map_y)), "r"(smem_u32(sblk)), "r"(n0 + c0), "r"(tile_w0),
                           "r"(tile_h0 + 4 * qd), "r"(n)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (STATS && !(p.dbg & 8)) {
          // lane l sums channels 2l, 2l+1 (one 32-bit word) over the warp's 32 pixel rows
          const uint8_t* sb = sblk + ((lane & 3) << 2);
          const int lch = lane >> 2;
          float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(sb + r * 128 + ((lch ^ (r & 7)) << 4));
            const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
            a0 += lo; a1 += hi;
            b0 = fmaf(lo, lo, b0); b1 = fmaf(hi, hi, b1);
          }
          __syncwarp();                                     // the next chunk overwrites the staging block
          st_s[c][0] += a0; st_s[c][1] += a1;
          st_q[c][0] += b0; st_q[c][1] += b1;
        }
      }
      if (p.dbg & 4) {                                      // (empty-epilogue probe: the loop above did not run)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader_cta) mbar_arrive(&bars->tmem_empty[buf]);
          else mbar_arrive_remote(empty_remote + (uint32_t)(buf * 8));
        }
      }
    }
    if (STATS && cur_n >= 0) flush(cur_n);
    if (STATS && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
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

// Second stage of the statistics: per (n, c) sum the slots of the CTAs that visited image n, in slot order, in fp64;
// then mean / rstd and the folded InstanceNorm + StyleMod affine (same outputs as sg_epi_finalize).
// A CTA (cluster jj, rank r) visits tiles 2q + r, q in [jj*Q/pairs, (jj+1)*Q/pairs); image n owns tiles [n*T, (n+1)*T).
__global__ void finalize_parts_kernel(const float* __restrict__ part, const float* __restrict__ style,
                                      float* __restrict__ mr, float* __restrict__ ab, int N, int C, int HW, int T, int Q,
                                      int pairs, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i % C;
  const int slots = 2 * pairs;
  double s1 = 0.0, s2 = 0.0;
  for (int s = 0; s < slots; ++s) {
    const int jj = s >> 1, r = s & 1;
    const int qb = (int)((long long)jj * Q / pairs), qe = (int)((long long)(jj + 1) * Q / pairs);
    if (qe <= qb) continue;
    int lo = max(n * T, 2 * qb + r);
    if ((lo - r) & 1) ++lo;
    const int hi = min((n + 1) * T - 1, 2 * (qe - 1) + r);
    if (lo > hi) continue;
    const float* q = part + (((size_t)n * slots + s) * 2) * C + c;
    s1 += (double)q[0];
    s2 += (double)q[C];
  }
  const double m = s1 / HW;
  double var = s2 / HW - m * m;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float mean = (float)m;
  mr[(size_t)i * 2] = mean;
  mr[(size_t)i * 2 + 1] = rstd;
  const float st0 = style ? style[(size_t)n * 2 * C + c] : 0.f;
  const float st1 = style ? style[(size_t)n * 2 * C + C + c] : 0.f;
  ab[(size_t)i * 3] = mean;
  ab[(size_t)i * 3 + 1] = rstd * (1.f + st0);
  ab[(size_t)i * 3 + 2] = st1;
}

// ---- host side ---------------------------------------------------------------------------------------------
static int pick_bn(int Cout) {
  if (Cout % 256 == 0) return 256;
  if (Cout == 128 || Cout == 64) return Cout;
  return 0;
}
static bool resident_ok(int Cin, int HB) { return 9 * (Cin / BK) * HB * BK * 2 <= 150 * 1024; }

struct Plan {
  int BN, HB, tiles_n, num_spatial, Q, pairs, kchunks, resident, w_bytes, stage_bytes, stages, bstages, tmem_cols, nacc, grid;
  size_t smem;
};

static bool make_plan(int N, int H, int W, int Cin, int Cout, Plan* pl, bool stats = true) {
  if (N < 1 || H % TH != 0 || W % TW != 0 || Cin % BK != 0 || Cin < BK) return false;
  const int BN = pick_bn(Cout);
  if (!BN) return false;
  pl->BN = BN; pl->HB = BN / 2;
  pl->tiles_n = Cout / BN;
  pl->num_spatial = N * (H / TH) * (W / TW);
  if (pl->num_spatial % 2 != 0) return false;
  pl->Q = pl->num_spatial / 2;
  pl->kchunks = Cin / BK;
  // 227 KB per CTA = static (bias / noise weights / reduction scratch 6 KB, + 32 KB statistics staging) + dynamic
  const int budget = 232448 - (stats ? 39040 : 6272) - (int)sizeof(Barriers) - 1024 - 256;
  // resident weights + as many activation slots as fit; else two rings: activation slots and (chunk, kw) weight slots
  const int b_slot = 3 * pl->HB * BK * 2;
  pl->stage_bytes = A_BYTES;
  pl->resident = 0; pl->bstages = 0;
  if (resident_ok(Cin, pl->HB)) {
    const int wb = 9 * pl->kchunks * pl->HB * BK * 2;
    int st = (budget - wb) / A_BYTES;
    if (st > MAX_STAGES) st = MAX_STAGES;
    if (st >= 2 && st >= (pl->kchunks < 3 ? pl->kchunks + 1 : 3)) { pl->resident = 1; pl->w_bytes = wb; pl->stages = st; }
  }
  if (!pl->resident) {
    int bst = 2, st = (budget - bst * b_slot) / A_BYTES;
    while (bst < MAX_BSTAGES && (budget - (bst + 1) * b_slot) / A_BYTES >= 4) {   // deepen the weight ring while >= 4 activation slots remain
      ++bst;
      st = (budget - bst * b_slot) / A_BYTES;
    }
    if (st > MAX_STAGES) st = MAX_STAGES;
    if (st < 2) return false;
    pl->stages = st; pl->bstages = bst; pl->w_bytes = bst * b_slot;
  }
  // all 512 TMEM columns (one CTA per SM): nacc = 512 / BN accumulator buffers (SG_PAIR_NACC=2: the first version)
  static const int nacc_env = getenv("SG_PAIR_NACC") ? atoi(getenv("SG_PAIR_NACC")) : 0;
  int nacc = 512 / BN;
  if (nacc > MAX_ACC) nacc = MAX_ACC;
  if (nacc_env >= 2 && nacc_env < nacc) nacc = nacc_env;
  if (nacc < 2) return false;
  pl->nacc = nacc;
  int cols = 32;
  while (cols < nacc * BN) cols <<= 1;
  pl->tmem_cols = cols;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int pairs = (sms / 2) / pl->tiles_n;                     // clusters per N block
  if (pairs > pl->Q) pairs = pl->Q;
  if (pairs < 1) return false;
  pl->pairs = pairs;
  pl->grid = pairs * pl->tiles_n * 2;
  pl->smem = (size_t)pl->stages * pl->stage_bytes + pl->w_bytes + sizeof(Barriers) + 1024;
  return true;
}

template <bool STATS, bool N64>
static int launch(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, const Params& p, const Plan& pl,
                  cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    err = cudaFuncSetAttribute(conv3x3_pair_kernel<STATS, N64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               232448 - (STATS ? 39040 : 6272));
  });
  if (err != cudaSuccess) { sg_set_error("sg_conv3x3_pair: cannot raise the dynamic shared memory limit"); return (int)err; }
  conv3x3_pair_kernel<STATS, N64><<<pl.grid, NUM_THREADS, pl.smem, st>>>(mx, mw, my, p);
  return sg_check_launch("sg_conv3x3_pair");
}

}  // namespace tc2

SG_EXPORT int sg_conv3x3_pair_supported(int N, int H, int W, int Cin, int Cout) {
  tc2::Plan pl;
  return tc2::make_plan(N, H, W, Cin, Cout, &pl) ? 1 : 0;
}

// number of statistic slots per image written by sg_conv3x3_pair (0 if unsupported)
SG_EXPORT int sg_conv3x3_pair_slots(int N, int H, int W, int Cin, int Cout) {
  tc2::Plan pl;
  return tc2::make_plan(N, H, W, Cin, Cout, &pl) ? 2 * pl.pairs : 0;
}

SG_EXPORT int sg_conv3x3_pair(const void* x, const void* Etc, void* y, const float* bias, float bias_mul, int act,
                              const float* noise, const float* noise_w, float* part, int N, int H, int W, int Cin,
                              int Cout, void* stream) {
  using namespace tc2;
  Plan pl;
  SG_REQUIRE(make_plan(N, H, W, Cin, Cout, &pl, part != nullptr), "sg_conv3x3_pair: shape not supported");
  EncodeTiledFn enc = get_encode();
  SG_REQUIRE(enc != nullptr, "sg_conv3x3_pair: cuTensorMapEncodeTiled unavailable");
  CUtensorMap map_x, map_w;
  SG_REQUIRE(encode_nhwc(&map_x, x, N, H, W, Cin, HALO_W, HALO_H), "sg_conv3x3_pair: cuTensorMapEncodeTiled(x) failed");
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)9 * Cout};
    cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)pl.HB};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(Etc), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SG_REQUIRE(r == CUDA_SUCCESS, "sg_conv3x3_pair: cuTensorMapEncodeTiled(w) failed");
  }
  Params p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.BN = pl.BN; p.HB = pl.HB;
  p.tiles_h = H / TH; p.tiles_w = W / TW; p.tiles_n = pl.tiles_n;
  p.num_spatial = pl.num_spatial; p.tiles_per_image = p.tiles_h * p.tiles_w;
  p.Q = pl.Q; p.pairs = pl.pairs;
  p.kchunks = pl.kchunks; p.stages = pl.stages; p.stage_bytes = pl.stage_bytes; p.bstages = pl.bstages; p.tmem_cols = pl.tmem_cols; p.nacc = pl.nacc;
  {
    static const int mmaw = getenv("SG_PAIR_MMAW") ? atoi(getenv("SG_PAIR_MMAW")) : MAX_MMA_WARPS;   // bench only
    p.mma_warps = mmaw < 1 ? 1 : (mmaw > MAX_MMA_WARPS ? MAX_MMA_WARPS : mmaw);
    // an issuer that starts a tile ahead waits on the SECOND fill of a ring slot when the ring is shorter than the tiles in
    // flight, and an mbarrier parity wait cannot tell 'two phases ahead' from 'done': one issuer then
    if (pl.stages < p.mma_warps * pl.kchunks || (!pl.resident && pl.bstages < p.mma_warps * 3 * pl.kchunks)) p.mma_warps = 1;
  }
  p.resident = pl.resident; p.w_bytes = pl.w_bytes;
  p.bias = bias; p.bias_mul = bias_mul; p.act = act;
  p.noise = noise; p.noise_w = noise_w; p.part = part;
  {
    static const int dbg = getenv("SG_PAIR_DBG") ? atoi(getenv("SG_PAIR_DBG")) : 0;
    p.dbg = dbg;
  }
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap map_y = map_x;                                // placeholder for the plain kernel (never dereferenced)
  if (part)                                                 // TMA-store boxes: 64 channels x 8 pixels x 4 rows (one warp's 32 pixels)
    SG_REQUIRE(encode_nhwc(&map_y, y, N, H, W, Cout, TW, 4), "sg_conv3x3_pair: cuTensorMapEncodeTiled(y) failed");
  static const bool no64 = getenv("SG_PAIR_NO64") != nullptr;            // bench only: the generic statistics epilogue at BN = 64
  if (part && pl.BN == 64 && !no64) return launch<true, true>(map_x, map_w, map_y, p, pl, st);
  return part ? launch<true, false>(map_x, map_w, map_y, p, pl, st) : launch<false, false>(map_x, map_w, map_y, p, pl, st);
}

// part as written by sg_conv3x3_pair for the same (N, H, W, Cin, Cout); style [N, 2*Cout] or null; mr [N,Cout,2];
// ab [N,Cout,3] = (mean, rstd*(1+s0), s1) -- the coefficient layout of sg_affine_nc.
SG_EXPORT int sg_epi_finalize_parts(const float* part, const float* style, float* mr, float* ab, int N, int H, int W,
                                    int Cin, int Cout, float eps, void* stream) {
  using namespace tc2;
  Plan pl;
  SG_REQUIRE(make_plan(N, H, W, Cin, Cout, &pl), "sg_epi_finalize_parts: shape not supported");
  const int T = (H / TH) * (W / TW);
  finalize_parts_kernel<<<sg_cdiv((long long)N * Cout, 128), 128, 0, (cudaStream_t)stream>>>(
      part, style, mr, ab, N, Cout, H * W, T, pl.Q, pl.pairs, eps);
  return sg_check_launch("sg_epi_finalize_parts");
}
