/* sgb200.h -- C ABI of libsgb200.so, the B200 (sm_100a) StyleGAN training-step kernels.
 *
 * The reference (huangzh13/StyleGAN.pytorch) has no FFI: its seam is the Python class API of models/*.py and all
 * arithmetic is delegated to torch.  Each entry point below replaces the torch call sites named
 * in its comment (paths relative to the reference root).  Conventions:
 *   - plain device pointers + sizes, no torch types; the caller owns every buffer (outputs, workspaces);
 *   - activations are NHWC ("channels-last") contiguous; `dtype` 0 = float32, 1 = bfloat16 (fp32 accumulation);
 *     parameters, gradients of parameters, statistics and images (NCHW) are always float32;
 *   - every call is asynchronous on `stream` (a cudaStream_t), never allocates, never synchronises;
 *   - returns 0 on success, otherwise a cudaError_t value or >=1000 for argument errors; sg_last_error() describes it.
 */
#ifndef SGB200_H
#define SGB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGB200_ABI_VERSION 2

/* Tap-table geometry shared by forward, dgrad and wgrad of every dense convolution (see csrc/sg_conv_generic.cu):
 *   out[n, qy*os+py, qx*os+px, :] = sum_t in[n, qy*is+dy[ph][t], qx*is+dx[ph][t], :] * E[ph][t]   (ph = 2*py+px) */
typedef struct SgConvGeom {
  int32_t N, Hq, Wq;
  int32_t Hin, Win, Cin;
  int32_t Hout, Wout, Cout;
  int32_t in_stride, out_stride;
  int32_t nphase, ntaps;
  int8_t dy[4][16];
  int8_t dx[4][16];
} SgConvGeom;

/* One row of the multi-tensor optimizer table (device memory).  Pointers are device addresses stored as int64. */
typedef struct SgTensorEntry {
  int64_t p, g, m, v, ema, step; /* ema may be 0; step -> float32 scalar (torch.optim.Adam state['step']) */
  int64_t n;                     /* elements */
  int64_t chunk0;                /* first chunk index of this tensor (chunks of sg_chunk_elems()) */
} SgTensorEntry;

const char* sg_last_error(void);
int sg_abi_version(void);
int sg_geom_size(void);
int sg_entry_size(void);
/* 1 (default): shape-specialised kernels where they apply (4-pixel blur strips, K=3 to/from-RGB, one-row weighted
 * channel sums); 0: general kernels only; <0: query.  Returns the previous setting.  Same results either way. */
int sg_fast_paths(int enable);

/* ---- grouped dense layers (csrc/sg_linear_group.cu): the 2 x layers StyleMod affines of one Generator forward,
 *      models/CustomLayers.py:203-216 (EqualizedLinear :82-103), as ONE launch each for forward, data gradient and
 *      weight/bias gradient.  `table` is device memory, G rows of sg_lingroup_rows() int64 words:
 *        { x ptr, ldx (elements), W ptr, bias ptr or 0, y ptr, O, scale | bias_mul << 32 (float bits), first block }
 *      forward:  y[B,O] = [lrelu](scale * x W^T + bias*bias_mul);   dgrad:  x-slot := [x-slot +] scale * y-slot W  (row
 *      stride ldx);   wgrad:  W-slot := scale * y-slot^T x,  bias-slot := bias_mul * column sums of y-slot.
 *      first block = prefix sum of sg_lingroup_fwd_blocks(O) (forward) or sg_lingroup_wgrad_blocks(O) (wgrad). */
int sg_lingroup_rows(void);
int sg_lingroup_supported(int B, int I);
int sg_lingroup_fwd_blocks(int O);
int sg_lingroup_wgrad_blocks(int O);
int sg_lingroup_fwd(const void* table, int G, int total_blocks, int B, int I, int act, void* stream);
int sg_lingroup_dgrad(const void* table, int G, int B, int I, int accumulate, void* stream);
int sg_lingroup_wgrad(const void* table, int G, int total_blocks, int B, int I, void* stream);

/* ---- dense convolutions: EqualizedConv2d.forward, models/CustomLayers.py:137-180 (F.conv2d :163,:171,:173,
 *      F.conv_transpose2d :151) and their autograd adjoints ------------------------------------------------------- */
/* E[ph][t][a][b] = scale * sum_j coef[ph*ntaps+t][j] * W[co][ci][j]; (a,b) = transpose ? (co,ci) : (ci,co).
 * Replaces `self.weight * self.w_mul` (:171) and the 4x4 kernel merges (:146-150, :159-162). */
int sg_wprep_fwd(const float* W, void* E, int dtype, int Cout, int Cin, int KK, int npt, const float* coef,
                 float scale, int transpose, void* stream);
/* the same for many weights in one launch (bf16 outputs): one table row per weight */
typedef struct SgWprepEntry {
  long long W, E, coef;          /* device pointers: float32 parameter [Cout][Cin][KK], bf16 output, float32 coef [npt][KK] */
  int Cout, Cin, KK, npt, transpose;
  int tile0;                     /* first block of this row: running sum of ceil(Cout/32)*ceil(Cin/32) */
  float scale;
  int pad;
} SgWprepEntry;
int sg_wprep_fwd_multi(const void* table, int n, int total_tiles, int max_smem_floats, void* stream);
/* dW = scale * coef^T . sum_splits(partial)  -- adjoint of sg_wprep_fwd applied to the wgrad partials. */
int sg_wprep_bwd(const float* partial, int nsplit, float* dW, int Cout, int Cin, int KK, int npt, const float* coef,
                 float scale, int transpose, int accumulate, void* stream);
int sg_conv_generic(const void* x, const void* E, void* y, const float* bias, float bias_mul, int act, int dtype,
                    const SgConvGeom* g, void* stream);
int sg_wgrad_splits(const SgConvGeom* g);
/* partial: [nsplit][nphase*ntaps][Cin][Cout] float32 */
int sg_wgrad_generic(const void* x, const void* dy, float* partial, int nsplit, int dtype, const SgConvGeom* g,
                     void* stream);

/* tcgen05/TMA implicit-GEMM 3x3 stride-1 convolution, bf16 operands, fp32 TMEM accumulation (csrc/sg_conv_tc.cu).
 * x [N,H,W,Cin] bf16, Etc [9][Cout][Cin] bf16 (K-major), y [N,H,W,Cout] bf16; fused bias + LeakyReLU epilogue.
 * The north-star kernel: blocks[i].conv1 / conv0, models/CustomLayers.py:170-171. */
int sg_conv3x3_tc_supported(int N, int H, int W, int Cin, int Cout);
int sg_conv3x3_tc(const void* x, const void* Etc, void* y, const float* bias, float bias_mul, int act, int N, int H,
                  int W, int Cin, int Cout, void* stream);

/* CTA-pair (tcgen05 cta_group::2, M = 256) form of the same convolution with the first half of LayerEpilogue fused into
 * the epilogue (csrc/sg_conv_pair_tc.cu): z = [lrelu](conv + bias*bias_mul + noise_w[c]*noise[n,h,w]) written once in
 * bf16, plus -- when `part` is given -- per-(image, CTA slot) partial sums of z and z^2 per channel for
 * nn.InstanceNorm2d (models/CustomLayers.py:191-200, :233 after :170-171).  noise fp32 [N,H,W] or null; part fp32
 * [N][sg_conv3x3_pair_slots][2][Cout] or null.  sg_epi_finalize_parts reduces the slots in fixed order (fp64) and
 * writes mr [N,Cout,2] = (mean, rstd) and ab [N,Cout,3] = (mean, rstd*(1+s0), s1) like sg_epi_finalize. */
int sg_conv3x3_pair_supported(int N, int H, int W, int Cin, int Cout);
int sg_conv3x3_pair_slots(int N, int H, int W, int Cin, int Cout);
int sg_conv3x3_pair(const void* x, const void* Etc, void* y, const float* bias, float bias_mul, int act,
                    const float* noise, const float* noise_w, float* part, int N, int H, int W, int Cin, int Cout,
                    void* stream);
int sg_epi_finalize_parts(const float* part, const float* style, float* mr, float* ab, int N, int H, int W, int Cin,
                          int Cout, float eps, void* stream);

/* BlurLayer + bias + the front half of LayerEpilogue in one pass (csrc/sg_blur_v2.cu; models/CustomLayers.py:174-180
 * then :242-246): z = lrelu(blur3(x) + bias*bias_mul + noise_w[c]*noise[n,h,w]) in bf16 and per-(image, block) partial
 * sums of z, z^2 (part [N][sg_blur3_epi_slots][2][C]); sg_epi_finalize_slots reduces them like sg_epi_finalize_parts. */
int sg_blur3_epi_slots(int N, int H, int W, int C);
int sg_blur3_epi(const void* x, void* z, const float* bias, float bias_mul, const float* noise, const float* noise_w,
                 float* part, int N, int H, int W, int C, void* stream);
int sg_epi_finalize_slots(const float* part, int slots, const float* style, float* mr, float* ab, int N, int C, int HW,
                          float eps, void* stream);

/* tcgen05/TMA up-/down-sampling convolutions in the tap-merged low-resolution form (csrc/sg_convud_tc.cu):
 * mode 0 UP   x [N,H,W,Cin] -> y [N,2H,2W,Cout], 4 phases x 2x2 taps  (CustomLayers.py:143-154 upscale conv; dgrad of down)
 * mode 1 DOWN x [N,2H,2W,Cin] -> y [N,H,W,Cout], 4x4 taps stride 2    (CustomLayers.py:156-180 downscale conv; dgrad of up)
 * E [16][Cout][Cin] bf16 from sg_wprep_fwd with the tables of models/_geom.py; fused bias + LeakyReLU. */
int sg_convud_tc_supported(int mode, int N, int H, int W, int Cin, int Cout);
/* kernel choice of sg_convud_tc (A/B tests): 2 (default) the halo-box CTA-pair kernel of csrc/sg_convud2_tc.cu where the
 * shape allows, 1 the first CTA-pair kernel, 0 the single-CTA kernel; < 0 queries.  Returns the previous value. */
int sg_pair_updown(int enable);
int sg_convud_tc(int mode, const void* x, const void* E, void* y, const float* bias, float bias_mul, int act, int N,
                 int H, int W, int Cin, int Cout, void* stream);

/* tcgen05/TMA weight gradient of the same convolution: part [S][9][Cout][Cin] float32 partial sums over S pixel splits
 * (S = sg_wgrad3x3_tc_splits); reduce + scale + OIHW layout with sg_wprep_bwd(transpose=1).  x [N,H,W,Cin], dy
 * [N,H,W,Cout] bf16.  Adjoint of models/CustomLayers.py:170-171 wrt the weight; also used by the R1 double backward. */
int sg_wgrad3x3_tc_supported(int N, int H, int W, int Cin, int Cout);
int sg_wgrad3x3_tc_splits(int N, int H, int W, int Cin, int Cout);
int sg_wgrad3x3_tc(const void* x, const void* dy, float* part, int N, int H, int W, int Cin, int Cout, void* stream);

/* Second-generation weight gradient (csrc/sg_wgrad2_tc.cu): every MMA is 128 x 192 x 16 -- N = the three vertical taps
 * of one 64-channel block of x (one halo box read as three 2048-byte-shifted MN-major blocks), M = two 64-channel
 * blocks of dy, or for Cout == 64 two horizontally shifted copies of dy.  Same `part` layout as sg_wgrad3x3_tc with
 * S = sg_wgrad3x3_v2_splits.  Covers Cin % 64 == 0 and Cout == 64 or Cout % 128 == 0. */
int sg_wgrad3x3_v2_supported(int N, int H, int W, int Cin, int Cout);
int sg_wgrad3x3_v2_splits(int N, int H, int W, int Cin, int Cout);
int sg_wgrad3x3_v2(const void* x, const void* dy, float* part, int N, int H, int W, int Cin, int Cout, void* stream);

/* tcgen05/TMA weight gradient of the up/down layers in the tap-merged form (csrc/sg_wgradud_tc.cu): part
 * [S][16][Cout][Cin] float32 (tap order of the layer's table), reduced onto the 3x3 parameter by sg_wprep_bwd with the
 * coefficient table of models/_geom.py (transpose=1).  H, W = low-resolution grid. */
int sg_wgradud_tc_supported(int mode, int N, int H, int W, int Cin, int Cout);
int sg_wgradud_tc_splits(int mode, int N, int H, int W, int Cin, int Cout);
int sg_wgradud_tc(int mode, const void* x, const void* dy, float* part, int N, int H, int W, int Cin, int Cout,
                  void* stream);

/* ---- bandwidth-bound feature-map kernels ------------------------------------------------------------------------ */
/* BlurLayer.forward, CustomLayers.py:266-276 (+ the `x + bias` of :178-179 when bias != NULL). Self-adjoint. */
int sg_blur3(const void* x, void* y, const float* bias, float bias_mul, int dtype, int N, int H, int W, int C,
             void* stream);
/* y = x + bias[c]*bias_mul + noise_w[c]*noise[n,p]: stand-alone NoiseLayer (CustomLayers.py:191-200) / bias add (:178). */
int sg_bias_noise(const void* x, const float* bias, float bias_mul, const float* noise, const float* noise_w, int act,
                  void* y, int dtype, long long P, int C, void* stream);
/* NHWC nearest x2 (Upscale2d, CustomLayers.py:27-36) and 2x2 box pooling with fused bias + LeakyReLU (Downscale2d,
 * CustomLayers.py:48-64, + :178-179): the pre/post passes of the tensor-core up/down convolutions. */
int sg_up2_nhwc(const void* x, void* y, float scale, int dtype, int N, int H, int W, int C, void* stream);
int sg_pool2_nhwc(const void* x, void* y, const float* bias, float bias_mul, int act, float scale, int dtype, int N,
                  int Ho, int Wo, int C, void* stream);
/* Zero-pad or crop an NHWC map to [N,hd,wd,C] (runs the 4x4/8x8 layers on the 8x16 tensor-core tiles). */
int sg_pad_crop_nhwc(const void* src, void* dst, int dtype, int N, int hs, int ws, int hd, int wd, int C,
                     void* stream);
/* out = a*x + b*y (y may be NULL): fade-in lerps GAN.py:202,427,586; scaling. */
int sg_axpby(const void* x, const void* y, void* out, float a, float b, long long n, int dtype, void* stream);
/* same with (a, b) = (ab[0], ab[1]) read from device memory: per-step scalars (the fade-in alpha, models/GAN.py:202,
 * 427, 588) must not be baked into a captured CUDA graph */
int sg_axpby_dev(const void* x, const void* y, void* out, const float* ab, long long n, int dtype, void* stream);
/* LeakyReLU(0.2) backward from the activation output (GAN.py:68,151,347). */
int sg_lrelu_bwd(const void* g, const void* ref, void* gx, long long n, int dtype, void* stream);
/* the same plus the bias gradient of the fused conv + bias + LeakyReLU it belongs to (CustomLayers.py:77-80 add + the
 * LeakyReLU of Blocks.py): gx over NHWC [P][C], gb[c] = scale * sum_p gx[p][c]; ws = sg_channel_sum_blocks(P) * C floats */
int sg_lrelu_bwd_bias_supported(int dtype, long long P, int C);
int sg_lrelu_bwd_bias(const void* g, const void* ref, void* gx, int dtype, long long P, int C, float* ws, float* gb,
                      float scale, void* stream);
int sg_cast(const void* x, int dtype_in, void* y, int dtype_out, long long n, void* stream);
/* out[j][c] = scale * sum_p wt_j(p) * x[p][c]; wt [N][J][HW] float32, optional all-ones last row.  Bias, noise-weight,
 * from_rgb and to_rgb weight gradients.  part_ws: sg_channel_sum_blocks(P) * rows * C floats. */
int sg_channel_sum_blocks(long long P);
int sg_channel_sum(const void* x, int dtype, const float* wt, int J, int ones, long long P, int HW, int C,
                   float* part_ws, float* out, float scale, int accumulate, void* stream);
/* from_rgb: 1x1 EqualizedConv2d on the NCHW float image + LeakyReLU, GAN.py:358,377,408,423-426 */
int sg_rgb_in(const float* img, const float* W, const float* bias, float scale, float bias_mul, int act, void* y,
              int dtype, int N, int HW, int K, int C, void* stream);
int sg_rgb_in_dgrad(const void* g, int dtype, const float* W, float scale, float* gimg, int N, int HW, int K, int C,
                    void* stream);
/* to_rgb (+ nearest x`up` of its input, + fade-in accumulate): GAN.py:157,167,173,199-202 */
int sg_rgb_out(const void* x, int dtype, const float* W, const float* bias, float scale, float bias_mul, float alpha,
               float beta, float* out, int N, int h, int w, int C, int K, int up, void* stream);
int sg_rgb_out_dgrad(const float* g, const float* W, float coef, void* gx, int dtype, int N, int h, int w, int C,
                     int K, int up, void* stream);
/* Input pipeline: RandomHorizontalFlip + ToTensor + Normalize(0.5, 0.5) of data/transforms.py:11-33 on a uint8 HWC batch
 * already on the device: in [N,H,W,K] bytes, flip [N] bytes (0/1, may be NULL), out [N,K,H,W] float32, bit-identical to
 * the CPU transform.  (csrc/sg_input.cu; not yet run on hardware) */
int sg_image_u8_to_f32(const unsigned char* in, const unsigned char* flip, float* out, int N, int H, int W, int K,
                       void* stream);
/* AvgPool2d / nearest interpolate on NCHW float images: GAN.py:382, 575-586 */
int sg_pool_nchw(const float* in, float* out, long long planes, int Ho, int Wo, int f, float scale, void* stream);
int sg_upsample_nchw(const float* in, float* out, long long planes, int Hi, int Wi, int f, float scale, void* stream);
/* ws: K*64 floats */
int sg_plane_sum(const float* g, float* out, float* ws, int N, int K, int HW, float scale, int accumulate,
                 void* stream);
/* style mixing + truncation: GAN.py:281-293, CustomLayers.py:319-323 */
int sg_dlatents(const float* w1, const float* w2, const float* avg, float* out, int B, int L, int D, int mix_cutoff,
                float psi, int trunc_cutoff, void* stream);
int sg_dlatents_bwd(const float* g, float* gw1, float* gw2, int B, int L, int D, int mix_cutoff, float psi,
                    int trunc_cutoff, int has_trunc, void* stream);
/* the style-mixing cut-off (host RNG draw, GAN.py:287-288) read from device memory, for CUDA-graph replay */
int sg_dlatents_dev(const float* w1, const float* w2, const float* avg, float* out, int B, int L, int D,
                    const int* mix_cutoff, float psi, int trunc_cutoff, void* stream);
int sg_dlatents_bwd_dev(const float* g, float* gw1, float* gw2, int B, int L, int D, const int* mix_cutoff, float psi,
                        int trunc_cutoff, int has_trunc, void* stream);

/* ---- LayerEpilogue: NoiseLayer + LeakyReLU + InstanceNorm2d + StyleMod, CustomLayers.py:191-248 ---------------- */
/* All reductions of this library are two-stage with a fixed summation order (no atomics): results are bitwise
 * reproducible from run to run.  `ws` = sg_epi_ws_floats(N, HW, C) floats of block partials; `sums` [N,C,3] doubles is
 * fully written by the call. */
long long sg_epi_ws_floats(int N, int HW, int C);
int sg_epi_act_stats(const void* y, const float* bias, float bias_mul, const float* noise, const float* noise_w,
                     int act, void* z, double* sums, float* ws, int dtype, int N, int HW, int C, void* stream);
int sg_epi_finalize(const double* sums, const float* style, float* mr, float* ab, int N, int C, int HW, float eps,
                    void* stream);
int sg_affine_nc(const void* z, const float* ab, void* out, int dtype, int N, int HW, int C, void* stream);
int sg_epi_bwd_reduce(const void* g, const void* z, const float* mr, double* sums, float* ws, int dtype, int N, int HW,
                      int C, void* stream);
int sg_epi_bwd_finalize(const double* sums, const float* mr, const float* style, float* gstyle, float* coef, int N,
                        int C, int HW, void* stream);
int sg_epi_bwd_apply(const void* g, const void* z, const float* coef, int act, void* gy, int dtype, int N, int HW,
                     int C, void* stream);

/* ---- small float32 kernels -------------------------------------------------------------------------------------- */
/* EqualizedLinear.forward, CustomLayers.py:99-103 (x rows `ldx` apart) */
/* wide inputs (I >= 2048: the discriminator's dense0) are split along I into sg_linear_fwd_ws_floats(B, I, O) floats
 * of partial products (0: no workspace needed), summed in slice order */
long long sg_linear_fwd_ws_floats(int B, int I, int O);
int sg_linear_fwd(const float* x, long long ldx, const float* W, const float* bias, float scale, float bias_mul,
                  int act, float* y, int B, int I, int O, float* ws, void* stream);
int sg_linear_dgrad(const float* gy, const float* W, float scale, float* gx, int B, int I, int O, void* stream);
int sg_linear_wgrad(const float* gy, const float* x, long long ldx, float scale, float bias_scale, float* gW,
                    float* gb, int B, int I, int O, int accumulate, void* stream);
/* PixelNormLayer, CustomLayers.py:22-23 */
int sg_pixelnorm_fwd(const float* x, float* y, int B, int D, float eps, void* stream);
int sg_pixelnorm_bwd(const float* g, const float* x, float* gx, int B, int D, float eps, void* stream);
/* StddevLayer, CustomLayers.py:294-305 (NHWC float32, ws >= 2*(B/G) floats) */
int sg_mbstd_fwd(const float* x, float* out, float* ws, int B, int HW, int C, int G, void* stream);
int sg_mbstd_bwd(const float* gout, const float* x, float* gx, float* ws, int B, int HW, int C, int G, void* stream);
int sg_mbstd_bwd2(const float* t, const float* gout, const float* x, float* cot_gout, float* cot_x, float* ws, int B,
                  int HW, int C, int G, void* stream);
/* LogisticGAN, Losses.py:213-229 */
int sg_softplus_sum(const float* x, int n, float sign, float scale, float* out, int accumulate, void* stream);
int sg_softplus_bwd(const float* x, int n, float sign, const float* gout, float gscale, float* gx, void* stream);
/* R1Penalty's sum of squares, Losses.py:209 (ws: 1024 doubles) */
int sg_sqsum(const float* x, long long n, float scale, float* out, int accumulate, double* ws, void* stream);
int sg_scale_by_scalar(const float* x, const float* gout, float coef, float* gx, long long n, void* stream);

/* ---- optimizer step: torch.optim.Adam + clip_grad_norm_ + update_average, GAN.py:529-533,651-656 ---------------- */
int sg_chunk_elems(void);
/* ws: total_chunks doubles (one partial per chunk, summed in chunk order by a second one-block kernel) */
int sg_sqnorm_multi(const SgTensorEntry* table_dev, int T, int total_chunks, double* out_dev, double* ws, void* stream);
int sg_adam_multi(const SgTensorEntry* table_dev, int T, int total_chunks, float lr, float beta1, float beta2,
                  float eps, float grad_scale, const double* normsq_dev, float max_norm, float ema_beta, void* stream);
int sg_ema_multi(const SgTensorEntry* table_dev, int T, int total_chunks, float ema_beta, void* stream);
/* NoiseLayer's torch.randn, CustomLayers.py:193 (Philox4x32-10 + Box-Muller) */
int sg_randn(float* out, long long n, unsigned long long seed, unsigned long long offset, void* stream);
int sg_randn_dev(float* out, long long n, unsigned long long seed, const unsigned long long* base,
                 unsigned long long delta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGB200_H */
