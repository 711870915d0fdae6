// Launchers of the StyleGAN2 glue kernels (sg2_kernels.cu). Device pointers, asynchronous on `st`.
#pragma once
#include "act_type.h"
#include <cuda_runtime.h>

namespace p2l {
typedef act_t bf16;

// y = act(wscale * x W^T + bias): WT is [in][out]; act 0 none, 1 leaky-relu(0.2)*sqrt2, 2 rsqrt(.+1e-8);
// square_in squares x first (demodulation)
void k_fc_fwd(const float* x, int ldx, const float* WT, const float* bias, float wscale, float* y, int ldy, int b, int in,
              int out, int act, int square_in, cudaStream_t st);
// same with an explicit row pitch of WT (a column block of a wider matrix)
void k_fc_fwd_ld(const float* x, int ldx, const float* WT, int ldw, const float* bias, float wscale, float* y, int ldy, int b,
                 int in, int out, int act, int square_in, cudaStream_t st);
// dx (+)= wscale * (dy * act'(y)) W, W is [out][in]
// splitk_scratch (>= 8 * 24 * in floats, or null): long reductions (out >= 2048) with few outputs are split over blocks
void k_fc_bwd(const float* dy, int lddy, const float* y, int ldy, const float* W, float wscale, float* dx, int lddx, int b,
              int in, int out, int act, int accumulate, cudaStream_t st, float* splitk_scratch = nullptr);
// the whole mapping network (PixelNorm + n_mlp EqualLinear(512, 512, fused_lrelu)) and its gradient as ONE cluster launch
// each (sdim == 512, b <= 24, n_mlp <= 8); h[0..n_mlp] are the [b, 512] activations, W^T / W as for k_fc_fwd / k_fc_bwd
bool k_sg_mapping_fusable(int b, int sdim, int n_mlp);
void k_sg_mapping_fwd(const float* z, const float* const* WT, const float* const* bias, float wscale, float* const* h, int b, int n_mlp,
                      cudaStream_t st);
void k_sg_mapping_bwd(const float* dw, const float* const* W, float wscale, float* const* h, const float* z, float* dz, float scale,
                      const float* row_scale, int b, int n_mlp, cudaStream_t st);
void k_pixelnorm_fwd(const float* x, float* y, int b, int n, cudaStream_t st);
void k_pixelnorm_bwd(const float* x, const float* dy, float* dx, int b, int n, float scale, const float* row_scale, cudaStream_t st);
void k_demod_bwd(const float* ddm, const float* dm, int lddm, const float* s, int lds, const float* Wsq, float* ds, int ldds,
                 int b, int Cin, int Cout, cudaStream_t st);
// A[b, p, c] = x[b or 0 (x_bstride 0), p, c] * s[b, c]: layer 0's modulated input (the learned constant times its style)
// Batched small GEMMs (one launch for all layers, b <= 24): host-side descriptor tables are filled with the *_desc calls,
// copied to the device once per plan, and launched with *_batched (max_J = the largest output width of the batch)
size_t k_sg_batch_bytes(int n);
void k_sg_demod_desc(void* host_tab, int i, const float* s, int lds, const float* wsqT, float* dm, int lddm, int b, int Cin, int Cout);
void k_sg_demod_bwd_desc(void* host_tab, int i, const float* ddm, const float* dm, int lddm, const float* s, int lds, const float* wsq,
                         float* ds, int ldds, int b, int Cin, int Cout);
void k_sg_demod_batched(const void* dev_tab, int n, int b, int max_J, cudaStream_t st);
void k_sg_demod_bwd_batched(const void* dev_tab, int n, int b, int max_J, cudaStream_t st);
void k_sg_modulate(const bf16* x, long x_bstride, const float* s, int lds, bf16* A, int b, int H, int W, int C, cudaStream_t st);
// The per-(sample, channel) reductions of the element-wise backward kernels (modulate_bwd -> ds, post_bwd_x -> ddm, torgb_bwd -> dweff)
// are two-stage and atomic-free: every 256-pixel block writes its partial sums to `scratch` (>= k_sg_scratch_floats
// floats), a second kernel adds them to the destination in block order.
long k_sg_scratch_floats(int b, int H, int W, int C);
void k_sg_modulate_bwd(const bf16* dA, const bf16* x, long x_bstride, const float* s, int lds, bf16* dx, float* ds, int ldds,
                       float* scratch, int b, int H, int W, int C, cudaStream_t st);
// The last layer's activation / noise / bias / demodulation backward (every other layer's lives in the next layer's
// dgrad epilogue, sg_epilogue.cuh): G = dm * dx * sqrt2 * lrelu'(x) ; ddm += sum_p g * u, u recomputed from x
void k_sg_post_bwd_x(const bf16* dx, const bf16* x, const float* dm, int lddm, const float* noise, const float* nw, const float* bias,
                     bf16* G, float* ddm, float* scratch, int b, int H, int W, int C, cudaStream_t st);
void k_sg_weff(const float* Wr, const float* s, int lds, float scale, float* weff, int b, int C, cudaStream_t st);
void k_sg_torgb_fwd(const bf16* x, const float* weff, const float* bias, const float* prev, float* rgb, int b, int H, int W, int C,
                    cudaStream_t st);
void k_sg_torgb_bwd(const float* drgb, const bf16* x, const float* weff, bf16* dx, float* dweff, float* scratch, int b, int H, int W,
                    int C, int accumulate, cudaStream_t st);
// last layer: ToRGB backward fused with that layer's activation / noise / bias / demodulation backward (k_sg_post_bwd_x)
void k_sg_torgb_post_bwd(const float* drgb, const bf16* x, const float* weff, float* dweff, const float* dm, int lddm, const float* noise,
                         const float* nw, const float* bias, bf16* G, float* ddm, float* scratch, int b, int H, int W, int C,
                         cudaStream_t st);
void k_sg_weff_bwd(const float* dweff, const float* Wr, float scale, float* ds, int ldds, int b, int C, cudaStream_t st);
void k_sg_rgb_up_adjoint(const float* drgb, float* dprev, int b, int h, int w, cudaStream_t st);
void k_sg_clamp(const float* rgb, float* img, long n, cudaStream_t st);
// NoiseInjection backward: dnoise[b,1,H,W] = nw * sum_c dx * sqrt2 * lrelu'(x), times scale (* row_scale[b])
void k_sg_noise_bwd(const bf16* dx, const bf16* x, const float* nw, float* dnoise, int b, int H, int W, int C, float scale,
                    const float* row_scale, cudaStream_t st);
void k_sg_scale_out(float* x, int b, long n, float scale, const float* row_scale, cudaStream_t st);
void k_sg_clamp_bwd(const float* rgb, const float* dimg, float* drgb, long n, float scale, cudaStream_t st);
}  // namespace p2l
