/* libp2l — C-ABI of the B200-native latent-inversion inner step.
 *
 * Drop-in boundary: in the reference the only caller of the generator and of the
 * loss is pix2latent/optimizer/closure.py:
 *     out  = model(**input_args)                              closure.py:51
 *     loss = loss_fn(out, **target_args).view(b,-1).mean(1)   closure.py:55
 *     loss.mean().backward()                                  closure.py:58
 * Every entry point below replaces one of those calls (cited per function). The reference has no
 * FFI of its own (it is pure Python on torch); the binding a maintainer adds is the ctypes stub
 * shown in INTEGRATION.md (pix2latent_b200/native.py is that stub).
 *
 * Conventions: plain C types; every pointer argument named *_dev is a DEVICE pointer owned by
 * the caller (torch); `stream` is a cudaStream_t passed as void*; every call is asynchronous and
 * stream-ordered (no hidden synchronisation); return 0 = ok, <0 = error with the message in
 * p2l_last_error() (thread-local). A handle is bound to one device and is not thread-safe.
 * There is no CPU fallback: p2l_create fails on anything that is not an sm_100 GPU.
 */
#ifndef P2L_H
#define P2L_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct p2l_ctx p2l_ctx;
typedef struct p2l_biggan p2l_biggan;
typedef struct p2l_lpips p2l_lpips;
typedef struct p2l_target p2l_target;
typedef struct p2l_sg2 p2l_sg2;

#define P2L_MAX_LAYERS 16

/* pytorch_pretrained_biggan/config.py (BigGANConfig) as used by pix2latent/model/biggan.py:26 */
typedef struct p2l_biggan_config {
    int n_layers;
    int up[P2L_MAX_LAYERS], in_mult[P2L_MAX_LAYERS], out_mult[P2L_MAX_LAYERS];
    int channel_width;      /* ch (128) */
    int z_dim;              /* 128 */
    int class_embed_dim;    /* 128 */
    int attention_pos;      /* SelfAttn inserted before block index attention_pos; <0: none */
    int n_stats;            /* 51 */
    float eps;              /* BN eps, 1e-4 */
    float truncation;       /* selects / interpolates the BN statistics row (biggan.py:50: 1.0) */
} p2l_biggan_config;

const char* p2l_last_error(void);
int p2l_version(void);
/* 16-bit storage / tensor-core operand type the library was built with: 0 = bfloat16, 1 = fp16 */
int p2l_act_dtype(void);
/* cumulative number of CUDA kernels this library has launched in this process */
long p2l_launch_count(void);

/* Bind to a CUDA device. Fails (-1) if the device is not compute capability 10.x. */
int p2l_create(int device, p2l_ctx** out);
void p2l_destroy(p2l_ctx* ctx);

/* ---- generator: replaces BigGAN.__init__ / forward (pix2latent/model/biggan.py:23-34, 50-58) */
int p2l_biggan_create(p2l_ctx* ctx, const p2l_biggan_config* cfg, p2l_biggan** out);
/* name = state-dict key of pix2latent's BigGAN module after remove_spectral_norm
 * (e.g. "generator.layers.0.conv_0.weight"); data = fp32, host or device, torch layout. */
int p2l_biggan_set_tensor(p2l_biggan* m, const char* name, const float* data, long numel);
/* Pack weights into the 16-bit (p2l_act_dtype) GEMM layouts the kernels read; frees the staging copies. */
int p2l_biggan_finalize(p2l_biggan* m);
void p2l_biggan_destroy(p2l_biggan* m);
/* img_dev[b,3,R,R] fp32 NCHW in (-1,1) = generator(cat(z,c), truncation). Keeps the activations
 * needed by p2l_biggan_backward for this batch size.                       biggan.py:58 */
int p2l_biggan_forward(p2l_biggan* m, int b, const float* z_dev, const float* c_dev, float* img_dev,
                       void* stream);
/* Given dL/dimg [b,3,R,R] of the LAST forward with this b: dz[b,z_dim], dc[b,class_embed_dim].
 * dgrad only — no weight gradients.                         closure.py:58 */
int p2l_biggan_backward(p2l_biggan* m, int b, const float* dimg_dev, float* dz_dev, float* dc_dev,
                        void* stream);
/* bytes of device memory held (weights + all cached per-batch plans) */
long p2l_biggan_device_bytes(p2l_biggan* m);
/* algorithmic FLOPs (MAC=2) of one forward for batch b through the tensor-core launches, and the
 * number of kernels launched by forward / backward (for bench.py's gpu_launches) */
double p2l_biggan_flops(p2l_biggan* m, int b, int backward);
int p2l_biggan_launches(p2l_biggan* m, int b, int backward);

/* ---- perceptual loss: replaces lpips.LPIPS(net, spatial=True) + ProjectionLoss
 *      (pix2latent/loss_functions.py:86-148) */
#define P2L_LPIPS_ALEX 0
#define P2L_LPIPS_VGG 1
int p2l_lpips_create(p2l_ctx* ctx, int net, p2l_lpips** out);
/* names: "net.slice{k}.{idx}.weight|bias" (torchvision feature indices) and "lin{k}.weight" */
int p2l_lpips_set_tensor(p2l_lpips* m, const char* name, const float* data, long numel);
int p2l_lpips_finalize(p2l_lpips* m);
void p2l_lpips_destroy(p2l_lpips* m);

/* Cache everything that depends only on the target: unit-normalised backbone features of
 * target[3,H,W], the adjoint-upsampled weight maps and sum(W). weight/mask may be NULL (ones).
 * rec_type 1 = l1, 2 = l2 (ReconstructionLoss); beta = LPIPS weight (ProjectionLoss: 10);
 * rec_weight = weight of the pixel term (1 for ProjectionLoss, 0 for PerceptualLoss alone),
 * per_weight likewise (0 for ReconstructionLoss alone).          loss_functions.py:97-100 */
int p2l_target_create(p2l_lpips* m, const float* target_dev, const float* weight_dev,
                      const float* mask_dev, int H, int W, int rec_type, float rec_weight,
                      float per_weight, p2l_target** out, void* stream);
void p2l_target_destroy(p2l_target* t);

/* loss_dev[b] = rec_weight*rec + per_weight*per for img_dev[b,3,H,W]. With want_grad, also
 * prepares d loss_i / d img_i (unit upstream) for p2l_loss_backward.    loss_functions.py:97 */
int p2l_loss_forward(p2l_lpips* m, p2l_target* t, int b, const float* img_dev, float* loss_dev,
                     int want_grad, void* stream);
/* dimg_dev[b,3,H,W] = dloss_dev[b] * d loss_b / d img_b  (of the last p2l_loss_forward) */
int p2l_loss_backward(p2l_lpips* m, p2l_target* t, int b, const float* dloss_dev, float* dimg_dev,
                      void* stream);
double p2l_lpips_flops(p2l_lpips* m, int b, int H, int W, int backward);
int p2l_lpips_launches(p2l_lpips* m, int backward);

/* ---- fused step: generator fwd + loss (+ backward to the latent) in one stream-ordered launch
 * sequence — what closure.py:51-58 does per mini-batch. The upstream gradient of sample i is
 * grad_scale * (dloss_dev ? dloss_dev[i] : 1); closure.py:58 (`loss.mean().backward()`) is
 * grad_scale = 1/b_chunk. loss_dev[b]; dz/dc/img may be NULL when not wanted. */
int p2l_biggan_step(p2l_biggan* g, p2l_lpips* l, p2l_target* t, int b, const float* z_dev,
                    const float* c_dev, int want_grad, float grad_scale, const float* dloss_dev,
                    float* loss_dev, float* dz_dev, float* dc_dev, float* img_dev, void* stream);

/* ---- transform search: per-candidate targets
 * p2l_affine_resample replaces SpatialTransform.transform / invert_transform
 * (pix2latent/transform/spatial_transform.py:69-104): dst[b,C,H,W] = F.grid_sample(src, F.affine_grid(theta,
 * src.size())) with torch's defaults (bilinear, zero padding, align_corners=False). theta_dev[b,2,3];
 * src_batch = b, or 1 when every row resamples the same source image. */
int p2l_affine_resample(const float* src_dev, int src_batch, const float* theta_dev, float* dst_dev, int b, int C,
                        int H, int W, void* stream);
/* p2l_biggan_step with one target per candidate (targets: HOST array of b handles created by
 * p2l_target_create with the same resolution / loss configuration; consecutive equal handles are batched):
 * what closure.py:51-58 computes when the 'target' / 'weight' variables were replaced per sample by
 * base_optimizer.py:61-79 (apply_transform). */
int p2l_biggan_step_targets(p2l_biggan* g, p2l_lpips* l, p2l_target* const* targets, int b, const float* z_dev,
                            const float* c_dev, int want_grad, float grad_scale, const float* dloss_dev,
                            float* loss_dev, float* dz_dev, float* dc_dev, float* img_dev, void* stream);

/* ---- device-resident inner loop: `steps` repetitions of closure.py:38-71 for one
 * population of b candidates without a host round trip —
 *     hooks: function_hooks.py:10-27 Clamp (clamp_z / clamp_c > 0: clamp to [-x, x] before every forward)
 *     out = model(z, c); loss = loss_fn(out, target...); loss.mean().backward()     closure.py:51-58
 *     optimizer.step(): torch.optim.Adam, one param group per latent tensor           closure.py:65,
 *                                                                            variable_manager.py:231-238
 * z_dev[b,z_dim] / c_dev[b,class_embed_dim] are updated IN PLACE. adam_mv_dev: [2][b*(z_dim+class_embed_dim)]
 * first/second moments laid out as (z rows | c rows), zero for a fresh optimizer; counters_dev: int[2],
 * [0] = Adam step count so far (carried across calls; 0 for a fresh optimizer), [1] = scratch.
 * loss_hist_dev[steps,b] receives every step's per-candidate losses (the values closure.step returns,
 * i.e. of the forward BEFORE that step's update). z_hist_dev / c_hist_dev (optional, [steps,b,dim]) record the
 * inputs as base_optimizer.py:105-106 tracks them (before the hooks of that step). img_dev (optional,
 * [b,3,R,R]) = output of the last forward. The upstream gradient of sample i is
 * grad_scale * (dloss_dev ? dloss_dev[i] : 1) as in p2l_biggan_step. use_graph: capture one step into a CUDA
 * graph and replay it (falls back to plain launches if capture is refused). Asynchronous, stream-ordered. */
typedef struct p2l_adam_config {
    float lr_z, lr_c;     /* learning rate of the z / c param groups (0 = frozen) */
    float beta1, beta2, eps;
    float clamp_z, clamp_c; /* Clamp hook bounds; <= 0: no hook */
} p2l_adam_config;
int p2l_biggan_optimize(p2l_biggan* g, p2l_lpips* l, p2l_target* t, int b, int steps, float* z_dev, float* c_dev,
                        const float* dloss_dev, float grad_scale, const p2l_adam_config* cfg, float* adam_mv_dev,
                        int* counters_dev, float* loss_hist_dev, float* z_hist_dev, float* c_hist_dev, float* img_dev,
                        int use_graph, void* stream);
/* 1 if the last p2l_biggan_optimize on this generator replayed a CUDA graph, 0 if it launched step by step */
int p2l_biggan_optimize_used_graph(p2l_biggan* g);
/* One Adam update alone (replaces `opt.step()` closure.py:65 for the latent leaves): same state layout as above;
 * loss_dev / loss_hist_dev may be NULL. */
int p2l_adam_update(int b, int z_dim, int c_dim, float* z_dev, float* c_dev, const float* dz_dev, const float* dc_dev,
                    const p2l_adam_config* cfg, float* adam_mv_dev, int* counters_dev, void* stream);

/* ---- StyleGAN2 generator: replaces StyleGAN2.__init__ / forward_z
 *      (pix2latent/model/stylegan2.py:66-119: rosinality Generator(size, 512, 8, channel_multiplier=2),
 *      `model([z], truncation=1.0)[0].clamp_(-1, 1)`) */
typedef struct p2l_sg2_config {
    int size;            /* output resolution (power of two, 8..1024) */
    int style_dim;       /* 512 */
    int n_mlp;           /* 8 */
    int channels[9];     /* feature channels at resolution 4, 8, ..., 1024 (multiples of 64) */
} p2l_sg2_config;
int p2l_sg2_create(p2l_ctx* ctx, const p2l_sg2_config* cfg, p2l_sg2** out);
/* name = key of rosinality's g_ema state dict ("style.1.weight", "conv1.conv.weight", "to_rgbs.0.bias", ...) */
int p2l_sg2_set_tensor(p2l_sg2* m, const char* name, const float* data, long numel);
int p2l_sg2_finalize(p2l_sg2* m);
void p2l_sg2_destroy(p2l_sg2* m);
int p2l_sg2_num_noise_layers(p2l_sg2* m);
/* img_dev[b,3,R,R] = clamp(G(z), -1, 1). noise_dev: HOST array of num_noise_layers DEVICE pointers to
 * the per-layer noise images [b,1,r,r] (r = 4,8,8,16,16,...); NULL = no noise. The caller draws the
 * noise (torch RNG), so reference runs can be replayed.      stylegan2.py:117-119 */
int p2l_sg2_forward(p2l_sg2* m, int b, const float* z_dev, const float* const* noise_dev, float* img_dev, void* stream);
int p2l_sg2_backward(p2l_sg2* m, int b, const float* dimg_dev, float* dz_dev, void* stream);
/* fused step, as p2l_biggan_step */
int p2l_sg2_step(p2l_sg2* g, p2l_lpips* l, p2l_target* t, int b, const float* z_dev, const float* const* noise_dev,
                 int want_grad, float grad_scale, const float* dloss_dev, float* loss_dev, float* dz_dev, float* img_dev,
                 void* stream);

/* ---- StyleGAN2 w / w+ / noise search: replaces StyleGAN2.forward_w
 *      (pix2latent/model/stylegan2.py:122-125: `model([w], input_is_latent=True, noise=noises)[0].clamp_(-1, 1)`)
 * latent_dev[b, n_latent, style_dim], n_latent = 2*log2(size) - 2: StyledConv l reads row l, ToRGB t reads row
 * 2t+1 (rosinality Generator.forward); a plain w is the same row repeated. The mapping network is skipped.
 * Backward reaches the latent rows and — when dnoise_dev (HOST array of num_noise_layers DEVICE pointers, entries
 * may be NULL) is given — every layer's noise image [b,1,r,r]. */
int p2l_sg2_n_latent(p2l_sg2* m);
/* w_dev[b, style_dim] = style(z) (PixelNorm + 8 EqualLinear): what stylegan2.py:97-104 samples 4096 times for
 * latent_mean / latent_std */
int p2l_sg2_style(p2l_sg2* m, int b, const float* z_dev, float* w_dev, void* stream);
int p2l_sg2_forward_w(p2l_sg2* m, int b, const float* latent_dev, const float* const* noise_dev, float* img_dev, void* stream);
int p2l_sg2_backward_w(p2l_sg2* m, int b, const float* dimg_dev, float* dlatent_dev, float* const* dnoise_dev, void* stream);
/* fused step, as p2l_sg2_step */
int p2l_sg2_step_w(p2l_sg2* g, p2l_lpips* l, p2l_target* t, int b, const float* latent_dev, const float* const* noise_dev,
                   int want_grad, float grad_scale, const float* dloss_dev, float* loss_dev, float* dlatent_dev,
                   float* const* dnoise_dev, float* img_dev, void* stream);

/* ---- measurement hooks (bench.py): while enabled, every tensor-core launch is bracketed by
 * CUDA events on its stream; p2l_profile_read synchronises and returns the summed duration
 * (ms), the number of launches and their algorithmic FLOPs since the last enable. */
void p2l_profile_enable(int on);
int p2l_profile_read(double* conv_ms, long* conv_launches, double* conv_flops);

#ifdef __cplusplus
}
#endif
#endif
