// Internal fast paths behind the C-ABI entry points of sg_pointwise.cu (not exported).  Each returns 0 when it has
// launched, SG_FAST_UNSUPPORTED when the shape is outside its domain (the entry point then runs the general kernel),
// or a cudaError_t.  sg_fast_paths(0) (C-ABI) switches them off so that tests can compare both implementations.
#pragma once
#define SG_FAST_UNSUPPORTED 1000
extern int g_sg_fast_paths;
int sgf_blur3(const void* x, void* y, const float* bias, float bias_mul, int dtype, int N, int H, int W, int C,
              void* stream);
int sgf_rgb_in3(const float* img, const float* W, const float* bias, float scale, float bias_mul, int act, void* y,
                int N, int HW, int C, void* stream);
int sgf_rgb_out3(const void* x, const float* W, const float* bias, float scale, float bias_mul, float alpha, float beta,
                 float* out, int N, int h, int w, int C, int up, void* stream);
int sgf_channel_sum_blocks(long long P);
int sgf_channel_sum_w3(const void* x, const float* wt, int ones, long long P, int HW, int C, float* ws, float* out,
                       float scale, void* stream);
int sgf_channel_sum_w1(const void* x, const float* wt, long long P, int C, float* ws, float* out, float scale,
                       void* stream);
int sgf_linear_fwd(const float* x, long long ldx, const float* W, const float* bias, float scale, float bias_mul, int act,
                   float* y, int B, int I, int O, void* stream);
int sgf_linear_wgrad(const float* gy, const float* x, long long ldx, float scale, float bias_scale, float* gW, float* gb,
                     int B, int I, int O, int accumulate, void* stream);
