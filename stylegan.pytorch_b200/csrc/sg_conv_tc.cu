// tcgen05 / TMA implicit-GEMM 3x3 stride-1 convolution for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
//   y[n,h,w,co] = act( sum_{kh,kw,ci} x[n,h+kh-1,w+kw-1,ci] * E[kh*3+kw][co][ci] + bias[co]*bias_mul )
//
// reference call site: EqualizedConv2d.forward -> F.conv2d(x, W*w_mul, bias, padding=1), models/CustomLayers.py:170-171
// (and, with flipped/transposed E, its dgrad).  No im2col: for each of the three horizontal shifts kw one TMA box
// {64 ch, TW px, TH+2 rows} of the NHWC activation lands in 128B-swizzled shared memory (out-of-bounds rows/columns
// are zero-filled by TMA = the conv padding); the three vertical taps kh are UMMA descriptors that start kh*TW rows
// into that same buffer, so every activation byte is fetched 3*(TH+2)/TH = 3.75x from L2 instead of 9x.
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-5 epilogue (TMEM -> regs -> bias,
// LeakyReLU -> bf16 -> global).  Persistent over output tiles of TH x TW = 128 pixels x BN channels, double-buffered
// TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "sg_common.cuh"
#include "../../include/sgb200.h"
#include "sg_tc_common.cuh"

namespace tc {

constexpr int TH = 8, TW = 16, BK = 64;
constexpr int HALO_ROWS = (TH + 2) * TW;           // 160 smem rows of 128 B per kw-shift
constexpr int A_BYTES = HALO_ROWS * BK * 2;        // 20480
constexpr int NUM_THREADS = 192;
constexpr int MAX_STAGES = 8;

struct __align__(8) Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t wfull;
  uint32_t tmem_base;
  uint32_t pad;
};

using namespace tcu;

struct Params {
  int N, H, W, Cin, Cout, BN;
  int tiles_h, tiles_w, tiles_n, num_spatial;
  int kchunks, stages, stage_bytes, tmem_cols;
  int resident, w_bytes;      // resident: all 9*kchunks weight tiles of this CTA's N block stay in shared memory
  const float* bias;
  float bias_mul;
  int act;
  __nv_bfloat16* y;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, Params p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzled tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem + (size_t)p.stages * p.stage_bytes;          // resident weights (if any)
  Barriers* bars = reinterpret_cast<Barriers*>(smem_w + p.w_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a CTA keeps one N block for its whole life and walks the spatial tiles with stride gridDim.x / tiles_n
  const int nt = blockIdx.x % p.tiles_n;
  const int sp0 = blockIdx.x / p.tiles_n, sp_step = gridDim.x / p.tiles_n;
  const int n0 = nt * p.BN;
  __shared__ float s_bias[128];
  for (int i = threadIdx.x; i < p.BN; i += blockDim.x) s_bias[i] = p.bias ? p.bias[n0 + i] * p.bias_mul : 0.f;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bars->tmem_full[a], 1); mbar_init(&bars->tmem_empty[a], 4); }
    mbar_init(&bars->wfull, 1);
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
  const int b_tap_bytes = p.BN * BK * 2;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (p.resident) {
        mbar_expect_tx(&bars->wfull, (uint32_t)p.w_bytes);
        for (int kc = 0; kc < p.kchunks; ++kc)
          for (int t = 0; t < 9; ++t)
            tma_load_2d(smem_w + (size_t)(kc * 9 + t) * b_tap_bytes, &map_w, &bars->wfull, kc * BK, t * p.Cout + n0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int sp = sp0; sp < p.num_spatial; sp += sp_step) {
        int r = sp;
        int wt = r % p.tiles_w; r /= p.tiles_w;
        int ht = r % p.tiles_h;
        int n = r / p.tiles_h;
        const int h0 = ht * TH, w0 = wt * TW;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int kw = 0; kw < 3; ++kw) {
            mbar_wait(&bars->empty[stage], phase ^ 1);
            uint8_t* sA = smem + (size_t)stage * p.stage_bytes;
            uint8_t* sB = sA + A_BYTES;
            mbar_expect_tx(&bars->full[stage], (uint32_t)(p.resident ? A_BYTES : A_BYTES + 3 * b_tap_bytes));
            tma_load_4d(sA, &map_x, &bars->full[stage], kc * BK, w0 + kw - 1, h0 - 1, n);
            if (!p.resident)
              for (int kh = 0; kh < 3; ++kh)
                tma_load_2d(sB + kh * b_tap_bytes, &map_w, &bars->full[stage], kc * BK, (kh * 3 + kw) * p.Cout + n0);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the pipeline with warp-uniform state; one elected lane issues the tcgen05 instructions.
    const uint32_t idesc = make_idesc_bf16(128, p.BN);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t sW = smem_u32(smem_w);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (p.resident) { mbar_wait(&bars->wfull, 0); tc_fence_after(); }
    const uint32_t b_kh_stride = p.resident ? (uint32_t)(3 * b_tap_bytes) : (uint32_t)b_tap_bytes;
    for (int sp = sp0; sp < p.num_spatial; sp += sp_step) {
      mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.BN);
      for (int it = 0; it < p.kchunks * 3; ++it) {
        mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        const uint32_t sA = smem_base + (uint32_t)(stage * p.stage_bytes);
        const int kw = it % 3, kc = it / 3;
        // weights: streamed with the stage [kh][BN][64] or resident [kc][kh*3+kw][BN][64]
        const uint32_t sB = p.resident ? (sW + (uint32_t)((kc * 9 + kw) * b_tap_bytes)) : (sA + A_BYTES);
        const uint64_t a0 = make_desc_sw128(sA);
        const uint64_t b0 = make_desc_sw128(sB);
        const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32);
        const uint32_t b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
        const uint32_t bstep = b_kh_stride >> 4;
        const bool leader = elect_one_sync();
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // descriptor start-address field is in 16-byte units: +kh*TW rows of 128 B, +k*32 B inside the atom
            const uint32_t al = a_lo + (uint32_t)((kh * (TW * 128) + k * 32) >> 4);
            const uint32_t bl = b_lo + kh * bstep + (uint32_t)((k * 32) >> 4);
            if (leader) umma_bf16_lohi(tmem_d, al, a_hi, bl, b_hi, idesc, (it | kh | k) ? 1u : 0u);
          }
        }
        if (leader) umma_commit(&bars->empty[stage]);          // frees the smem stage once these MMAs have read it
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one_sync()) umma_commit(&bars->tmem_full[acc]);  // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue warps (TMEM lane quadrant = warp % 4) =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int pr = m / TW, pc = m % TW;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int sp = sp0; sp < p.num_spatial; sp += sp_step) {
      int r = sp;
      int wt = r % p.tiles_w; r /= p.tiles_w;
      int ht = r % p.tiles_h;
      int n = r / p.tiles_h;
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      __nv_bfloat16* dst = p.y + ((((size_t)n * p.H + (ht * TH + pr)) * p.W + (wt * TW + pc)) * p.Cout + n0);
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
        // up to four 16-column TMEM loads in flight, one wait
        uint32_t v0[16], v1[16], v2[16], v3[16];
        const int nld = (p.BN - c0) >= 64 ? 4 : (p.BN - c0) / 16;
        tmem_ld16_nowait(taddr + c0, v0);
        if (nld > 1) tmem_ld16_nowait(taddr + c0 + 16, v1);
        if (nld > 2) tmem_ld16_nowait(taddr + c0 + 32, v2);
        if (nld > 3) tmem_ld16_nowait(taddr + c0 + 48, v3);
        tmem_ld_wait();
        if (c0 + 64 >= p.BN) {                              // last chunk is in registers: release the TMEM buffer now
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

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
// N block: 128 whenever Cout allows it, otherwise the whole Cout (16..112).  A 128xNx16 SS-mode MMA costs ~61 cycles
// plus ~0.15 cycles per column (measured: N=64 67.9, N=128 79.9, N=256 130.4 cycles, tools/mma_rate.cu), so the
// work per instruction should be as large as the accumulator budget allows.
static int pick_bn(int Cin, int Cout) {
  (void)Cin;
  if (Cout % 128 == 0) return 128;
  if (Cout <= 128 && Cout % 16 == 0) return Cout;
  return 0;
}
// Resident-weight mode: the 9 * ceil(Cin/64) weight tiles of one N block fit beside >= 3 activation stages.
static bool resident_ok(int Cin, int BN) { return 9 * ((Cin + BK - 1) / BK) * BN * BK * 2 <= 150 * 1024; }

}  // namespace tc

SG_EXPORT int sg_conv3x3_tc_supported(int N, int H, int W, int Cin, int Cout) {
  if (N < 1 || H % tc::TH != 0 || W % tc::TW != 0) return 0;
  // Cin need not be a multiple of 64: the TMA boxes are 64 channels wide and the part beyond the tensor is
  // zero-filled in shared memory (only the 16-byte global stride rule applies)
  if (Cin < 16 || Cin % 8 != 0) return 0;
  if (tc::pick_bn(Cin, Cout) == 0) return 0;
  return 1;
}

SG_EXPORT int sg_conv3x3_tc(const void* x, const void* Etc, void* y, const float* bias, float bias_mul, int act, int N,
                            int H, int W, int Cin, int Cout, void* stream) {
  using namespace tc;
  SG_REQUIRE(sg_conv3x3_tc_supported(N, H, W, Cin, Cout), "sg_conv3x3_tc: shape not supported");
  EncodeTiledFn enc = get_encode();
  SG_REQUIRE(enc != nullptr, "sg_conv3x3_tc: cuTensorMapEncodeTiled unavailable");
  const int BN = pick_bn(Cin, Cout);

  CUtensorMap map_x, map_w;
  SG_REQUIRE(encode_nhwc(&map_x, x, N, H, W, Cin, TW, TH + 2), "sg_conv3x3_tc: cuTensorMapEncodeTiled(x) failed");
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)9 * Cout};
    cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(Etc), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SG_REQUIRE(r == CUDA_SUCCESS, "sg_conv3x3_tc: cuTensorMapEncodeTiled(w) failed");
  }

  Params p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.BN = BN;
  p.tiles_h = H / TH; p.tiles_w = W / TW; p.tiles_n = Cout / BN;
  p.num_spatial = N * p.tiles_h * p.tiles_w;
  p.kchunks = (Cin + BK - 1) / BK;
  p.resident = resident_ok(Cin, BN) ? 1 : 0;
  p.w_bytes = p.resident ? 9 * p.kchunks * BN * BK * 2 : 0;
  p.stage_bytes = p.resident ? A_BYTES : A_BYTES + 3 * BN * BK * 2;
  int stages = (216 * 1024 - p.w_bytes) / p.stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  SG_REQUIRE(stages >= 2, "sg_conv3x3_tc: not enough shared memory for two stages");
  p.stages = stages;
  int cols = 2 * BN;
  int tc_cols = 32;
  while (tc_cols < cols) tc_cols <<= 1;
  p.tmem_cols = tc_cols;
  p.bias = bias; p.bias_mul = bias_mul; p.act = act;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);

  size_t smem = (size_t)stages * p.stage_bytes + p.w_bytes + sizeof(Barriers) + 1024;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  });
  SG_REQUIRE(attr_err == cudaSuccess, "sg_conv3x3_tc: cannot raise the dynamic shared memory limit");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = (sms / p.tiles_n) * p.tiles_n;                  // every CTA keeps one N block
  if (grid > p.num_spatial * p.tiles_n) grid = p.num_spatial * p.tiles_n;
  SG_REQUIRE(grid >= p.tiles_n, "sg_conv3x3_tc: more N blocks than SMs");
  conv3x3_tc_kernel<<<grid, NUM_THREADS, smem, (cudaStream_t)stream>>>(map_x, map_w, p);
  return sg_check_launch("sg_conv3x3_tc");
}
