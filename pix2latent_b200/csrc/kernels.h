// Launchers for the non-contraction kernels of the inversion step (memory-bound glue around
// the tcgen05 convolutions). All pointers are device pointers; all launches are asynchronous
// on `st`. bf16 tensors are NHWC.
#pragma once
#include "act_type.h"
#include <cuda_runtime.h>

namespace p2l {

typedef act_t bf16;  // historical alias: the 16-bit activation type (fp16 by default, act_type.h)

// ---- latent side ---------------------------------------------------------------------------
// cond[b,256] = cat(z[b,zd], c[b,cd])
void k_concat_cond(const float* z, const float* c, float* cond, int b, int zd, int cd, cudaStream_t st);
// a[b, ch] = (1 + cond.Ws[ch]) * inv_std[ch] ; s[b, ch] = cond.Wo[ch] - mean[ch]*a   (conditional BNs), as one
// GEMV + bias: WT_as = [cdim][2*C_cond] = [Ws*inv_std | Wo - mean*inv_std*Ws] (transposed), bias_as =
// [inv_std | -mean*inv_std] — the BN statistics are folded in when the weights are packed
void k_cond_affine(const float* cond, const float* WT_as, const float* bias_as, float* a, float* s, int b, int cdim,
                   int C_cond, int stride, cudaStream_t st);
// rows [C_cond, C_cond+C_unc): a = weight*inv_std, s = bias - mean*a (same for every sample)
void k_uncond_affine(const float* weight, const float* bias, const float* mean, const float* inv_std,
                     float* a, float* s, int b, int C_cond, int C_unc, int stride, cudaStream_t st);
// h[b, j] = cond[b].W[j] + bias[j] ; raw = bf16(h) ; act = relu(a[b, j%C]*h + s[b, j%C])
// W is stored TRANSPOSED: [cdim][J]
void k_gen_z(const float* cond, const float* W, const float* bias, const float* a, const float* s,
             int aff_stride, bf16* raw, bf16* act, int b, int cdim, int J, int C, cudaStream_t st);
// BN-affine gradient finalisation: from S0 = sum dpre, S1 = sum dpre*y to
//   G[b, ch] = dgamma = inv_std*(da - mean*ds),  G[b, C+ch] = dbeta = ds,
//   with da = (S1 - s*S0)/a, ds = S0.
void k_bn_grad_finalize(const float* S0, const float* S1, const float* a, const float* s,
                        const float* mean, const float* inv_std, float* G, int b, int C_cond,
                        int stride, cudaStream_t st);
// partial[blk, b, k] = sum_{j in block blk} G[b, j] * W[j, k]   (W row-major [J, cdim], G row stride ldg;
// k_dcond_blocks(J) blocks; cdim % 128 == 0)
int k_dcond_blocks(int J);
void k_dcond_partial(const float* G, int ldg, const float* W, float* partial, int b, int J, int cdim, cudaStream_t st);
// dcond = sum of nblk partial blocks; dz/dc = its halves * scale * (row_scale ? row_scale[b] : 1)
void k_dcond_reduce_split(const float* partial, int nblk, float* dz, float* dc, int b, int zd, int cd, float scale,
                          const float* row_scale, cudaStream_t st);
// G32[b, j] = float(g_bf16[b, j])
void k_bf16_to_f32(const bf16* src, float* dst, long n, cudaStream_t st);
// dz/dc = dcond halves * scale * (row_scale ? row_scale[b] : 1)
void k_split_dcond(const float* dcond, float* dz, float* dc, int b, int zd, int cd, float scale,
                   const float* row_scale, cudaStream_t st);

// ---- BigGAN glue ---------------------------------------------------------------------------
// rgb head, second half: T[n, tap*3+o, y, x] holds every pixel's contribution to its 3x3 neighbours
// (x[n,y,x,:] . W[o,:,tap]); img[n,o,y,x] = tanh(bias[o] + sum_tap T[n, tap*3+o, y+r-1, x+s-1]) with zero padding
void k_rgb_gather(const float* T, const float* bias, float* img, int b, int H, int W, cudaStream_t st);
// backward through nearest-x2 + relu + BN affine of an up block's conv_0 output:
//   g = sum2x2(g_up) ; dpre = g*[y>0] ; dx = a*dpre ; the sums of dpre and dpre*y over each 256-pixel block go to
//   statp[((n * parts + block) * 2 + {0,1}) * C + c], parts = k_pool_bnrelu_parts(H, W) (no atomics; k_stat_reduce sums them)
int k_pool_bnrelu_parts(int H, int W);
void k_pool_bnrelu_bwd(const bf16* g_up, const bf16* y_lo, const float* a, int aff_stride,
                       float* statp, bf16* dx, int b, int H, int W, int C, cudaStream_t st);
// One 32-channel chunk of one BN layer's partial sums (ConvGemmParams::statp layout): S0/S1[n][off + c0 + lane] =
// sum over `parts` slots, in slot order
struct StatSeg { const float* p; int parts, pstride, C, off, c0, off1; };  // parts filled of pstride allocated per image;
                                                                           // off / off1: column offsets into S0 / S1
void k_stat_reduce(const StatSeg* segs, int nsegs, float* S0, float* S1, int stride, int b, cudaStream_t st);
// the two sums go to different tables: S0[n * stride0 + off + c], S1[n * stride1 + off1 + c] (StyleGAN2: ds, ddm)
void k_stat_reduce2(const StatSeg* segs, int nsegs, float* S0, int stride0, float* S1, int stride1, int b, cudaStream_t st);

// out[b,H,W,C] = sum of the 2x2 block of in[b,2H,2W,inC] (first C channels): the skip gradient of an
// up block at the block's input resolution
void k_pool2x2_sum(const bf16* in, int inC, bf16* out, int b, int H, int W, int C, cudaStream_t st);

// ---- attention glue ------------------------------------------------------------------------
// 2x2 max-pool of channels [c0, c0+C) of x[b,H,W,xC] -> out[b,(H/2)*(W/2),C], outT[b,C,(H/2)*(W/2)]
// (either may be null), argmax (0..3) -> idx
void k_maxpool2_fwd(const bf16* x, int xC, int c0, int C, bf16* out, bf16* outT, unsigned char* idx,
                    int b, int H, int W, cudaStream_t st);
// dx[b,H,W,xC] channels [c0,c0+C) = routed d_out[b,(H/2)(W/2),C]
void k_maxpool2_bwd(const bf16* d_out, const unsigned char* idx, bf16* dx, int xC, int c0, int C,
                    int b, int H, int W, cudaStream_t st);
// P = softmax(S) row-wise, S fp32 [rows, n] -> P bf16
void k_softmax_fwd(const float* S, bf16* P, long rows, int n, cudaStream_t st);
// dS = P * (dP - sum_k dP*P)
void k_softmax_bwd(const bf16* P, const float* dP, bf16* dS, long rows, int n, cudaStream_t st);
// D[p] = sum_c a[p, c] * b[p, c]  (softmax backward: rowsum(dP o P) = dO . O per query row)
void k_rowdot(const bf16* a, const bf16* b, float* D, long rows, int C, cudaStream_t st);
// out[b, c, r] = in[b, r, in_c0 + c], in row stride ldin
void k_transpose(const bf16* in, int ldin, int in_c0, bf16* out, int b, int R, int C, cudaStream_t st);

// ---- image / loss glue ---------------------------------------------------------------------
// AlexNet conv1 (11x11 s4 p2) im2col of the scaled image: col[b,Ho,Wo,Kp] (k = (c*11+r)*11+s, zero pad)
void k_im2col_alex1(const float* img, bf16* col, int b, int H, int W, int Ho, int Wo, int Kp, cudaStream_t st);
// gradient back to the (unscaled) image: dimg[b,3,H,W] (+)= col2im(dcol) / scale_c
void k_col2im_alex1(const bf16* dcol, float* dimg, int b, int H, int W, int Ho, int Wo, int Kp, int accumulate, float unscale, cudaStream_t st);
// VGG first layer: img fp32 NCHW -> scaled bf16 NHWC padded to Cp channels
void k_img_to_nhwc_scaled(const float* img, bf16* out, int b, int H, int W, int Cp, cudaStream_t st);
void k_nhwc_to_dimg_scaled(const bf16* dx, int Cp, float* dimg, int b, int H, int W, int accumulate, float unscale, cudaStream_t st);
// max-pool k x k stride s (no padding), with argmax byte
void k_maxpool_fwd(const bf16* x, bf16* out, unsigned char* idx, int b, int H, int W, int C, int Ho, int Wo,
                   int k, int s, cudaStream_t st);
// dx = [x>0]*(sum over windows whose argmax is this pixel of dout) + addin
void k_maxpool_bwd(const bf16* dout, const unsigned char* idx, const bf16* x, const bf16* addin, bf16* dx,
                   int b, int H, int W, int C, int Ho, int Wo, int k, int s, cudaStream_t st);
// LPIPS layer distance + its gradient: f[b,HW,C] bf16 (post-relu), t[HW,C] fp32 (unit-normalised
// target features), lin[C], wadj[HW] (adjoint-upsampled weight map / sumW * beta).
//   sum_p wadj[p] * sum_c lin_c (f_c/(|f|+eps) - t_c)^2 ; g = d/df (masked by f>0), or null.
// Deterministic reduction: block k of sample bi writes its part to lossp[bi * lp_stride + k] (k_lpips_dist_slots(HW)
// slots); k_loss_reduce sums a sample's slots in slot order.
int k_lpips_dist_slots(int HW);
void k_lpips_dist(const bf16* f, const float* t, const float* lin, const float* wadj, float* lossp, int lp_stride,
                  bf16* g, int b, int HW, int C, float gscale, cudaStream_t st);
void k_loss_reduce(const float* lossp, int nslots, int lp_stride, float* loss, int b, cudaStream_t st);
// t[p, c] = f_c/(|f|+eps)  (target features, b = 1)
void k_lpips_normalize(const bf16* f, float* t, int HW, int C, cudaStream_t st);
// wadj_k = U_k^T (sum_c W[c]) * coef   for a feature map h x w (bilinear, align_corners=False)
void k_upsample_adjoint(const float* wsum, float* wadj, int H, int W, int h, int w, float coef, cudaStream_t st);
// wsum[p] = sum_c W[c,p] (* mask) ; total = sum
void k_weight_sum(const float* weight, const float* mask, float* wsum, float* total, int HW, cudaStream_t st);
// L1 term: sum |t-o| * W / sumW into k_l1_loss_slots(HW3) partial slots per sample (see k_lpips_dist) ;
// dimg[b,c,p] = -sign(t-o) * W / sumW   (dimg overwritten)
int k_l1_loss_slots(int HW3);
void k_l1_loss(const float* img, const float* target, const float* weight, const float* mask,
               const float* total, float* lossp, int lp_stride, float* dimg, int b, int HW3, int HW, int l2, cudaStream_t st);
// dimg[b] *= dloss[b]
void k_scale_rows(float* x, const float* scale, int b, long n, cudaStream_t st);
// rgb conv backward input: A[b,H,W,Kp] bf16 with k = (r*3+s)*3 + c of dpre = dimg*(1-img^2)
// `scale` = gradient scale applied on entry to 16-bit storage (act_type.h kGradScale)
void k_im2col_rgb_bwd(const float* dimg, const float* img, bf16* col, int b, int H, int W, int Kp, float scale, cudaStream_t st);

void k_fill_f32(float* p, float v, long n, cudaStream_t st);

// ---- transform search (pix2latent/transform/spatial_transform.py:69-104) ---------------------
// dst[b,C,H,W] = grid_sample(src, affine_grid(theta[b,2,3], src.size())) with torch's defaults (bilinear,
// zeros padding, align_corners=False); src_batch 1: every output row samples the same source image
void k_affine_resample(const float* src, int src_batch, const float* theta, float* dst, int b, int C, int H, int W,
                       cudaStream_t st);

}  // namespace p2l
