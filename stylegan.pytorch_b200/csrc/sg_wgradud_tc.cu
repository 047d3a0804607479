// tcgen05 / TMA weight gradient of the up- and down-sampling convolutions in their tap-merged form.
//
//   UP   layer:  part[s][ph*4 + a*2 + b][co][ci] = sum_{i,j} dy[2i+py, 2j+px][co] * x[i+a+py-1, j+b+px-1][ci]
//   DOWN layer:  part[s][u*4 + v][co][ci]        = sum_{i,j} dy[i, j][co]        * x[2i-1+u, 2j-1+v][ci]
// (i,j) runs over the LOW-resolution grid; sg_wprep_bwd folds the 16 effective taps back onto the 3x3 parameter with
// the transposed coefficient table of models/_geom.py (the adjoint of the merge in CustomLayers.py:146-150,159-162).
//
// Same machinery as sg_wgrad_tc.cu (MN-major operands straight from NHWC, M = Cout block, N = Cin block, K = pixels).
// A CTA owns (g, M block, N block, pixel split) with g = output phase (UP) or horizontal tap v (DOWN); per 8x16 tile it
// runs two stages, one per x box (UP: b = 0,1; DOWN: row parity a = 0,1), each box serving two vertical taps ->
// four accumulators of BN columns in TMEM.  The stride-2 operand (dy for UP, x for DOWN) is fetched by TMA with
// element strides 2, so no up-sampled or gathered copy of anything is ever materialised.
#include "sg_common.cuh"
#include "../../include/sgb200.h"
#include "sg_tc_common.cuh"

namespace wgud {

using namespace tcu;

constexpr int TH = 8, TW = 16;
constexpr int DY_BLOCK_BYTES = TH * TW * 128;          // 16384
constexpr int X_BLOCK_BYTES = (TH + 1) * TW * 128;     // 18432
constexpr int NUM_THREADS = 192;
constexpr int MAX_STAGES = 8;
constexpr int MODE_UP = 0;                        // mode 1 = DOWN (include/sgb200.h)

struct __align__(8) Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t done;
  uint32_t tmem_base;
  uint32_t pad;
};

struct Params {
  int mode, N, H, W, Cin, Cout, BM, BN;     // H, W: low-resolution grid
  int tiles_h, tiles_w, total_tiles;
  int MB, NB, S;
  int stages, stage_bytes, tmem_cols;
  float* part;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgradud_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + (size_t)p.stages * p.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int u_ = blockIdx.x;
  const int nb = u_ % p.NB; u_ /= p.NB;
  const int mb = u_ % p.MB; u_ /= p.MB;
  const int g = u_ % 4;                       // UP: phase (py,px);  DOWN: horizontal tap v
  const int s = u_ / 4;
  const int per = (p.total_tiles + p.S - 1) / p.S;
  const int t0 = s * per;
  const int t1 = (t0 + per < p.total_tiles) ? (t0 + per) : p.total_tiles;
  const int a_blocks = p.BM / 64, b_blocks = p.BN / 64;
  const int a_bytes = a_blocks * DY_BLOCK_BYTES;
  const int py = g >> 1, px = g & 1;

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
        for (int box = 0; box < 2; ++box) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sA = smem + (size_t)stage * p.stage_bytes;
          uint8_t* sB = sA + a_bytes;
          mbar_expect_tx(&bars->full[stage], (uint32_t)(a_bytes + b_blocks * X_BLOCK_BYTES));
          int dyw, dyh, xw, xh;
          if (p.mode == MODE_UP) {            // dy strided at the phase offset; x box b = box
            dyw = 2 * w0 + px; dyh = 2 * h0 + py;
            xw = w0 + box + px - 1; xh = h0 + py - 1;
          } else {                            // dy plain; x strided, row parity a = box, horizontal tap v = g
            dyw = w0; dyh = h0;
            xw = 2 * w0 + g - 1; xh = 2 * h0 + box - 1;
          }
          for (int j = 0; j < a_blocks; ++j)
            tma_load_4d(sA + j * DY_BLOCK_BYTES, &map_dy, &bars->full[stage], mb * p.BM + 64 * j, dyw, dyh, n);
          for (int j = 0; j < b_blocks; ++j)
            tma_load_4d(sB + j * X_BLOCK_BYTES, &map_x, &bars->full[stage], nb * p.BN + 64 * j, xw, xh, n);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // The GEMM N dimension is the TWO vertical taps of one 64-channel block of x: the (TH+1)-row box is an MN-major operand
    // whose 64-element blocks are 2048 bytes (one 16-pixel row) apart, so blocks 0/1 of the same buffer are the taps
    // 0/1 -> N = 128 in one instruction instead of two N = 64 ones (80 instead of 2 x 68 cycles, profiles/r01_mma_rate.txt).
    const uint32_t idesc = make_idesc_bf16_mn(p.BM, 128);
    const uint32_t smem_base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = t0; tile < t1; ++tile) {
      const uint32_t not_first = (tile != t0) ? 1u : 0u;
#pragma unroll
      for (int box = 0; box < 2; ++box) {
        mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        const uint32_t sA = smem_base + (uint32_t)(stage * p.stage_bytes);
        const uint32_t sB = sA + (uint32_t)a_bytes;
        const uint64_t a0 = make_desc_mn_sw128(sA, DY_BLOCK_BYTES);
        const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32);
        const bool leader = elect_one_sync();
        for (int j = 0; j < b_blocks; ++j) {                // accumulator (box, j): columns [t][64 channels of block j]
          const uint64_t b0 = make_desc_mn_sw128(sB + (uint32_t)(j * X_BLOCK_BYTES), TW * 128);
          const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
#pragma unroll
          for (int k = 0; k < (TH * TW) / 16; ++k) {
            const uint32_t ko = (uint32_t)((k * 2048) >> 4);
            if (leader)
              umma_bf16_lohi(tmem_base + (uint32_t)((box * b_blocks + j) * 128), a_lo + ko, a_hi, b_lo + ko, b_hi, idesc,
                             k ? 1u : not_first);
          }
        }
        if (leader) umma_commit(&bars->empty[stage]);
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    if (elect_one_sync()) umma_commit(&bars->done);
    __syncwarp();
  } else {
    const int q = warp & 3;
    mbar_wait(&bars->done, 0);
    tc_fence_after();
    const bool m64 = (p.BM == 64);
    const int row = m64 ? (q * 16 + lane) : (q * 32 + lane);
    const bool active = m64 ? (lane < 16) : true;
    const int co = mb * p.BM + row;
    for (int box = 0; box < 2; ++box) {
      for (int t = 0; t < 2; ++t) {
        // effective tap index in the layer's table order
        const int widx = (p.mode == MODE_UP) ? (g * 4 + t * 2 + box)          // [phase][a = t][b = box]
                                             : ((box + 2 * t) * 4 + g);       // [u = a + 2*bb][v = g], a = box, bb = t
        float* dst = p.part + ((((size_t)s * 16 + widx) * p.Cout + co) * p.Cin + (size_t)nb * p.BN);
        for (int c0 = 0; c0 < p.BN; c0 += 16) {
          uint32_t v[16];
          // accumulator (box, x block j = c0 / 64): tap t occupies its columns [t*64, t*64+64)
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) +
                                 (uint32_t)((box * b_blocks + (c0 >> 6)) * 128 + t * 64 + (c0 & 63));
          tmem_ld16(taddr, v);
          if (active && co < p.Cout && nb * p.BN + c0 < p.Cin) {
            float4* d4 = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              d4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                  __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          }
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

static bool encode_nhwc_s(CUtensorMap* map, const void* ptr, int N, int H, int W, int C, int box_w, int box_h,
                          int stride) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)(stride * box_w), (cuuint32_t)(stride * box_h), 1u};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace wgud

// H, W: low-resolution grid.  mode 0 (UP layer): x [N,H,W,Cin], dy [N,2H,2W,Cout];  mode 1 (DOWN layer): x [N,2H,2W,Cin],
// dy [N,H,W,Cout].
SG_EXPORT int sg_wgradud_tc_supported(int mode, int N, int H, int W, int Cin, int Cout) {
  if (mode != 0 && mode != 1) return 0;
  if (N < 1 || H % wgud::TH != 0 || W % wgud::TW != 0) return 0;
  if (wgud::pick_block(Cin) == 0 || wgud::pick_block(Cout) == 0) return 0;
  return 1;
}

SG_EXPORT int sg_wgradud_tc_splits(int mode, int N, int H, int W, int Cin, int Cout) {
  if (!sg_wgradud_tc_supported(mode, N, H, W, Cin, Cout)) return 0;
  int total = N * (H / wgud::TH) * (W / wgud::TW);
  int units = 4 * ((Cout + wgud::pick_block(Cout) - 1) / wgud::pick_block(Cout)) * ((Cin + wgud::pick_block(Cin) - 1) / wgud::pick_block(Cin));
  int S = (148 + units - 1) / units;
  if (S > total) S = total;
  if (S < 1) S = 1;
  int per = (total + S - 1) / S;
  S = (total + per - 1) / per;
  return S;
}

// part: [S][16][Cout][Cin] float32
SG_EXPORT int sg_wgradud_tc(int mode, const void* x, const void* dy, float* part, int N, int H, int W, int Cin,
                            int Cout, void* stream) {
  using namespace wgud;
  SG_REQUIRE(sg_wgradud_tc_supported(mode, N, H, W, Cin, Cout), "sg_wgradud_tc: shape not supported");
  CUtensorMap map_dy, map_x;
  if (mode == MODE_UP) {
    SG_REQUIRE(encode_nhwc_s(&map_dy, dy, N, 2 * H, 2 * W, Cout, TW, TH, 2), "sg_wgradud_tc: tensor map (dy) failed");
    SG_REQUIRE(encode_nhwc_s(&map_x, x, N, H, W, Cin, TW, TH + 1, 1), "sg_wgradud_tc: tensor map (x) failed");
  } else {
    SG_REQUIRE(encode_nhwc_s(&map_dy, dy, N, H, W, Cout, TW, TH, 1), "sg_wgradud_tc: tensor map (dy) failed");
    SG_REQUIRE(encode_nhwc_s(&map_x, x, N, 2 * H, 2 * W, Cin, TW, TH + 1, 2), "sg_wgradud_tc: tensor map (x) failed");
  }
  Params p;
  p.mode = mode; p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.BM = pick_block(Cout); p.BN = pick_block(Cin);
  p.tiles_h = H / TH; p.tiles_w = W / TW; p.total_tiles = N * p.tiles_h * p.tiles_w;
  p.MB = (Cout + p.BM - 1) / p.BM; p.NB = (Cin + p.BN - 1) / p.BN;
  p.S = sg_wgradud_tc_splits(mode, N, H, W, Cin, Cout);
  p.stage_bytes = (p.BM / 64) * DY_BLOCK_BYTES + (p.BN / 64) * X_BLOCK_BYTES;
  int stages = (214 * 1024) / p.stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  SG_REQUIRE(stages >= 2, "sg_wgradud_tc: not enough shared memory");
  p.stages = stages;
  int cols = 4 * p.BN, tc_cols = 32;
  while (tc_cols < cols) tc_cols <<= 1;
  p.tmem_cols = tc_cols;
  p.part = part;
  size_t smem = (size_t)stages * p.stage_bytes + sizeof(Barriers) + 1024;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(wgradud_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  SG_REQUIRE(attr_err == cudaSuccess, "sg_wgradud_tc: cannot raise the dynamic shared memory limit");
  int grid = p.S * 4 * p.MB * p.NB;
  wgradud_tc_kernel<<<grid, NUM_THREADS, smem, (cudaStream_t)stream>>>(map_dy, map_x, p);
  return sg_check_launch("sg_wgradud_tc");
}
