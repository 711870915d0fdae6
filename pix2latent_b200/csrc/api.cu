// C-ABI entry points (include/p2l.h). Thin: argument checks, handle casts, error strings.
#include "p2l.h"

#include "biggan.h"
#include "lpips.h"
#include "sg2.h"
#include "optim.h"

#define P2L_EXPORT extern "C" __attribute__((visibility("default")))

using namespace p2l;

struct p2l_ctx { Ctx c; };
struct p2l_biggan { BigGAN g; InnerLoop loop; };
struct p2l_lpips { Lpips l; };
struct p2l_target { Target* t; };
struct p2l_sg2 { SG2 g; };

#define P2L_TRY_BEGIN try {
#define P2L_TRY_END                                             \
    }                                                           \
    catch (const std::exception& e) {                           \
        set_error("exception: %s", e.what());                   \
        return -1;                                              \
    }

P2L_EXPORT int p2l_version(void) { return 1; }
P2L_EXPORT int p2l_act_dtype(void) { return P2L_ACT_FP16 ? 1 : 0; }
P2L_EXPORT long p2l_launch_count(void) { return launch_count(); }

P2L_EXPORT int p2l_create(int device, p2l_ctx** out) {
    if (!out) { set_error("p2l_create: out is NULL"); return -1; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        set_error("p2l_create: no CUDA device visible (this library has no CPU fallback)");
        return -1;
    }
    if (device < 0 || device >= n) { set_error("p2l_create: device %d out of range (%d devices)", device, n); return -1; }
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
    if (major != 10) {
        set_error("p2l_create: device %d is sm_%d%d; this library is built for sm_100a only", device, major, minor);
        return -1;
    }
    P2L_CUDA_CHECK(cudaSetDevice(device));
    p2l_ctx* c = new p2l_ctx();
    c->c.device = device;
    cudaDeviceGetAttribute(&c->c.sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = c;
    return 0;
}
P2L_EXPORT void p2l_destroy(p2l_ctx* ctx) { delete ctx; }

// ----------------------------------------------------------------------------- BigGAN
P2L_EXPORT int p2l_biggan_create(p2l_ctx* ctx, const p2l_biggan_config* cfg, p2l_biggan** out) {
    if (!ctx || !cfg || !out) { set_error("p2l_biggan_create: NULL argument"); return -1; }
    P2L_TRY_BEGIN
    p2l_biggan* m = new p2l_biggan();
    m->g.ctx = &ctx->c;
    m->g.cfg = *cfg;
    *out = m;
    return 0;
    P2L_TRY_END
}
P2L_EXPORT int p2l_biggan_set_tensor(p2l_biggan* m, const char* name, const float* data, long numel) {
    if (!m || !name || !data) { set_error("p2l_biggan_set_tensor: NULL argument"); return -1; }
    if (m->g.finalized) { set_error("p2l_biggan_set_tensor after finalize"); return -1; }
    P2L_TRY_BEGIN
    return m->g.stage.set(name, data, numel);
    P2L_TRY_END
}
P2L_EXPORT int p2l_biggan_finalize(p2l_biggan* m) {
    if (!m) { set_error("p2l_biggan_finalize: NULL"); return -1; }
    P2L_TRY_BEGIN
    return m->g.finalize();
    P2L_TRY_END
}
P2L_EXPORT void p2l_biggan_destroy(p2l_biggan* m) { delete m; }
P2L_EXPORT int p2l_biggan_forward(p2l_biggan* m, int b, const float* z, const float* c, float* img, void* stream) {
    if (!m || b <= 0 || !z || !c) { set_error("p2l_biggan_forward: bad argument"); return -1; }
    P2L_TRY_BEGIN
    return m->g.forward(b, z, c, img, static_cast<cudaStream_t>(stream));
    P2L_TRY_END
}
P2L_EXPORT int p2l_biggan_backward(p2l_biggan* m, int b, const float* dimg, float* dz, float* dc, void* stream) {
    if (!m || b <= 0 || !dimg || !dz || !dc) { set_error("p2l_biggan_backward: bad argument"); return -1; }
    P2L_TRY_BEGIN
    return m->g.backward(b, dimg, dz, dc, static_cast<cudaStream_t>(stream));
    P2L_TRY_END
}
P2L_EXPORT long p2l_biggan_device_bytes(p2l_biggan* m) {
    if (!m) return 0;
    return (long)m->g.device_bytes();
}
P2L_EXPORT double p2l_biggan_flops(p2l_biggan* m, int b, int backward) {
    if (!m) return 0;
    return m->g.flops(b, backward);
}
P2L_EXPORT int p2l_biggan_launches(p2l_biggan* m, int b, int backward) {
    (void)m; (void)b; (void)backward;
    return 0;  // superseded by p2l_launch_count(); kept for ABI stability
}

// ----------------------------------------------------------------------------- LPIPS
P2L_EXPORT int p2l_lpips_create(p2l_ctx* ctx, int net, p2l_lpips** out) {
    if (!ctx || !out) { set_error("p2l_lpips_create: NULL argument"); return -1; }
    if (net != P2L_LPIPS_ALEX && net != P2L_LPIPS_VGG) { set_error("p2l_lpips_create: unknown net %d", net); return -1; }
    P2L_TRY_BEGIN
    p2l_lpips* m = new p2l_lpips();
    m->l.ctx = &ctx->c;
    m->l.net = net;
    *out = m;
    return 0;
    P2L_TRY_END
}
P2L_EXPORT int p2l_lpips_set_tensor(p2l_lpips* m, const char* name, const float* data, long numel) {
    if (!m || !name || !data) { set_error("p2l_lpips_set_tensor: NULL argument"); return -1; }
    if (m->l.finalized) { set_error("p2l_lpips_set_tensor after finalize"); return -1; }
    P2L_TRY_BEGIN
    return m->l.stage.set(name, data, numel);
    P2L_TRY_END
}
P2L_EXPORT int p2l_lpips_finalize(p2l_lpips* m) {
    if (!m) { set_error("p2l_lpips_finalize: NULL"); return -1; }
    P2L_TRY_BEGIN
    return m->l.finalize();
    P2L_TRY_END
}
P2L_EXPORT void p2l_lpips_destroy(p2l_lpips* m) { delete m; }

P2L_EXPORT int p2l_target_create(p2l_lpips* m, const float* target, const float* weight, const float* mask, int H, int W,
                                 int rec_type, float rec_weight, float per_weight, p2l_target** out, void* stream) {
    if (!m || !target || !out || H <= 0 || W <= 0) { set_error("p2l_target_create: bad argument"); return -1; }
    if (rec_type != 1 && rec_type != 2) { set_error("p2l_target_create: rec_type must be 1 (l1) or 2 (l2)"); return -1; }
    P2L_TRY_BEGIN
    Target* t = m->l.make_target(target, weight, mask, H, W, rec_type, rec_weight, per_weight, static_cast<cudaStream_t>(stream));
    if (!t) return -1;
    p2l_target* h = new p2l_target();
    h->t = t;
    *out = h;
    return 0;
    P2L_TRY_END
}
P2L_EXPORT void p2l_target_destroy(p2l_target* t) {
    if (t) { delete t->t; delete t; }
}
P2L_EXPORT int p2l_loss_forward(p2l_lpips* m, p2l_target* t, int b, const float* img, float* loss, int want_grad, void* stream) {
    if (!m || !t || b <= 0 || !img || !loss) { set_error("p2l_loss_forward: bad argument"); return -1; }
    P2L_TRY_BEGIN
    return m->l.loss_forward(*t->t, b, img, loss, want_grad, static_cast<cudaStream_t>(stream));
    P2L_TRY_END
}
P2L_EXPORT int p2l_loss_backward(p2l_lpips* m, p2l_target* t, int b, const float* dloss, float* dimg, void* stream) {
    if (!m || !t || b <= 0 || !dloss || !dimg) { set_error("p2l_loss_backward: bad argument"); return -1; }
    P2L_TRY_BEGIN
    return m->l.loss_backward(*t->t, b, dloss, dimg, static_cast<cudaStream_t>(stream));
    P2L_TRY_END
}
P2L_EXPORT double p2l_lpips_flops(p2l_lpips* m, int b, int H, int W, int backward) {
    if (!m) return 0;
    return m->l.flops(b, H, W, backward);
}
P2L_EXPORT int p2l_lpips_launches(p2l_lpips* m, int backward) {
    (void)m; (void)backward;
    return 0;
}

P2L_EXPORT void p2l_profile_enable(int on) { profile_enable(on); }
P2L_EXPORT int p2l_profile_read(double* conv_ms, long* conv_launches, double* conv_flops) {
    return profile_read(conv_ms, conv_launches, conv_flops);
}

// ----------------------------------------------------------------------------- fused step
P2L_EXPORT int p2l_biggan_step(p2l_biggan* g, p2l_lpips* l, p2l_target* t, int b, const float* z, const float* c,
                               int want_grad, float grad_scale, const float* dloss, float* loss, float* dz, float* dc,
                               float* img, void* stream) {
    if (!g || !l || !t || b <= 0 || !z || !c || !loss) { set_error("p2l_biggan_step: bad argument"); return -1; }
    if (want_grad && (!dz || !dc)) { set_error("p2l_biggan_step: want_grad needs dz and dc"); return -1; }
    P2L_TRY_BEGIN
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (g->g.forward(b, z, c, img, st)) return -1;
    const float* im = img ? img : g->g.last_image(b);
    if (l->l.loss_forward(*t->t, b, im, loss, want_grad, st)) return -1;
    if (!want_grad) return 0;
    float* dimg = l->l.unit_grad(*t->t, b);
    if (!dimg) return -1;
    if (g->g.backward(b, dimg, dz, dc, st, grad_scale, dloss)) return -1;
    return 0;
    P2L_TRY_END
}

// ----------------------------------------------------------------------------- transform search
P2L_EXPORT int p2l_affine_resample(const float* src, int src_batch, const float* theta, float* dst, int b, int C, int H, int W,
                                   void* stream) {
    if (!src || !theta || !dst || b <= 0 || C <= 0 || H <= 0 || W <= 0 || (src_batch != 1 && src_batch != b)) {
        set_error("p2l_affine_resample: bad argument");
        return -1;
    }
    P2L_TRY_BEGIN
    k_affine_resample(src, src_batch, theta, dst, b, C, H, W, static_cast<cudaStream_t>(stream));
    P2L_CUDA_CHECK(cudaGetLastError());
    return 0;
    P2L_TRY_END
}
P2L_EXPORT int p2l_biggan_step_targets(p2l_biggan* g, p2l_lpips* l, p2l_target* const* targets, int b, const float* z,
                                       const float* c, int want_grad, float grad_scale, const float* dloss, float* loss,
                                       float* dz, float* dc, float* img, void* stream) {
    if (!g || !l || !targets || b <= 0 || !z || !c || !loss) { set_error("p2l_biggan_step_targets: bad argument"); return -1; }
    if (want_grad && (!dz || !dc)) { set_error("p2l_biggan_step_targets: want_grad needs dz and dc"); return -1; }
    P2L_TRY_BEGIN
    std::vector<Target*> Ts((size_t)b);
    for (int i = 0; i < b; ++i) {
        if (!targets[i] || !targets[i]->t) { set_error("p2l_biggan_step_targets: target %d is NULL", i); return -1; }
        Ts[i] = targets[i]->t;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (g->g.forward(b, z, c, img, st)) return -1;
    const float* im = img ? img : g->g.last_image(b);
    if (l->l.loss_forward_multi(Ts.data(), b, im, loss, want_grad, st)) return -1;
    if (!want_grad) return 0;
    float* dimg = l->l.unit_grad(*Ts[0], b);
    if (!dimg) return -1;
    if (g->g.backward(b, dimg, dz, dc, st, grad_scale, dloss)) return -1;
    return 0;
    P2L_TRY_END
}

// ----------------------------------------------------------------------------- device-resident inner loop
P2L_EXPORT int p2l_biggan_optimize(p2l_biggan* g, p2l_lpips* l, p2l_target* t, int b, int steps, float* z, float* c,
                                   const float* dloss, float grad_scale, const p2l_adam_config* cfg, float* mv, int* counters,
                                   float* loss_hist, float* z_hist, float* c_hist, float* img, int use_graph, void* stream) {
    if (!g || !l || !t || b <= 0 || steps < 0 || !z || !c || !cfg || !mv || !counters) {
        set_error("p2l_biggan_optimize: bad argument");
        return -1;
    }
    if (!(cfg->beta1 >= 0.f && cfg->beta1 < 1.f && cfg->beta2 >= 0.f && cfg->beta2 < 1.f && cfg->eps >= 0.f)) {
        set_error("p2l_biggan_optimize: bad Adam hyper-parameters (beta1 %g beta2 %g eps %g)", cfg->beta1, cfg->beta2, cfg->eps);
        return -1;
    }
    P2L_TRY_BEGIN
    return biggan_optimize(g->g, l->l, *t->t, g->loop, b, steps, z, c, dloss, grad_scale, *cfg, mv, counters, loss_hist, z_hist,
                           c_hist, img, use_graph, static_cast<cudaStream_t>(stream));
    P2L_TRY_END
}
P2L_EXPORT int p2l_biggan_optimize_used_graph(p2l_biggan* g) { return g ? g->loop.graph_used : 0; }
P2L_EXPORT int p2l_adam_update(int b, int z_dim, int c_dim, float* z, float* c, const float* dz, const float* dc,
                               const p2l_adam_config* cfg, float* mv, int* counters, void* stream) {
    if (b <= 0 || z_dim < 0 || c_dim < 0 || !cfg || !mv || !counters || (z_dim > 0 && (!z || !dz)) || (c_dim > 0 && (!c || !dc))) {
        set_error("p2l_adam_update: bad argument");
        return -1;
    }
    P2L_TRY_BEGIN
    const int nz = b * z_dim, nc = b * c_dim;
    k_adam(z, c, dz, dc, mv, mv + nz + nc, nz, nc, *cfg, counters, nullptr, nullptr, 0, static_cast<cudaStream_t>(stream));
    P2L_CUDA_CHECK(cudaGetLastError());
    return 0;
    P2L_TRY_END
}

// ----------------------------------------------------------------------------- StyleGAN2
P2L_EXPORT int p2l_sg2_create(p2l_ctx* ctx, const p2l_sg2_config* cfg, p2l_sg2** out) {
    if (!ctx || !cfg || !out) { set_error("p2l_sg2_create: NULL argument"); return -1; }
    P2L_TRY_BEGIN
    p2l_sg2* m = new p2l_sg2();
    m->g.ctx = &ctx->c;
    m->g.cfg = *cfg;
    *out = m;
    return 0;
    P2L_TRY_END
}
P2L_EXPORT int p2l_sg2_set_tensor(p2l_sg2* m, const char* name, const float* data, long numel) {
    if (!m || !name || !data) { set_error("p2l_sg2_set_tensor: NULL argument"); return -1; }
    if (m->g.finalized) { set_error("p2l_sg2_set_tensor after finalize"); return -1; }
    P2L_TRY_BEGIN
    return m->g.stage.set(name, data, numel);
    P2L_TRY_END
}
P2L_EXPORT int p2l_sg2_finalize(p2l_sg2* m) {
    if (!m) { set_error("p2l_sg2_finalize: NULL"); return -1; }
    P2L_TRY_BEGIN
    return m->g.finalize();
    P2L_TRY_END
}
P2L_EXPORT void p2l_sg2_destroy(p2l_sg2* m) { delete m; }
P2L_EXPORT int p2l_sg2_num_noise_layers(p2l_sg2* m) { return m ? m->g.num_layers : 0; }
P2L_EXPORT int p2l_sg2_forward(p2l_sg2* m, int b, const float* z, const float* const* noise, float* img, void* stream) {
    if (!m || b <= 0 || !z) { set_error("p2l_sg2_forward: bad argument"); return -1; }
    P2L_TRY_BEGIN
    return m->g.forward(b, z, noise, img, static_cast<cudaStream_t>(stream));
    P2L_TRY_END
}
P2L_EXPORT int p2l_sg2_backward(p2l_sg2* m, int b, const float* dimg, float* dz, void* stream) {
    if (!m || b <= 0 || !dimg || !dz) { set_error("p2l_sg2_backward: bad argument"); return -1; }
    P2L_TRY_BEGIN
    return m->g.backward(b, dimg, dz, static_cast<cudaStream_t>(stream));
    P2L_TRY_END
}
P2L_EXPORT int p2l_sg2_step(p2l_sg2* g, p2l_lpips* l, p2l_target* t, int b, const float* z, const float* const* noise,
                            int want_grad, float grad_scale, const float* dloss, float* loss, float* dz, float* img, void* stream) {
    if (!g || !l || !t || b <= 0 || !z || !loss) { set_error("p2l_sg2_step: bad argument"); return -1; }
    if (want_grad && !dz) { set_error("p2l_sg2_step: want_grad needs dz"); return -1; }
    P2L_TRY_BEGIN
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (g->g.forward(b, z, noise, img, st)) return -1;
    const float* im = img ? img : g->g.last_image(b);
    if (l->l.loss_forward(*t->t, b, im, loss, want_grad, st)) return -1;
    if (!want_grad) return 0;
    float* dimg = l->l.unit_grad(*t->t, b);
    if (!dimg) return -1;
    return g->g.backward(b, dimg, dz, st, grad_scale, dloss);
    P2L_TRY_END
}

// ----------------------------------------------------------------------------- StyleGAN2 w / w+ / noise search
P2L_EXPORT int p2l_sg2_n_latent(p2l_sg2* m) { return (m && m->g.finalized) ? m->g.n_latent() : 0; }
P2L_EXPORT int p2l_sg2_style(p2l_sg2* m, int b, const float* z, float* w, void* stream) {
    if (!m || b <= 0 || !z || !w) { set_error("p2l_sg2_style: bad argument"); return -1; }
    P2L_TRY_BEGIN
    return m->g.style(b, z, w, static_cast<cudaStream_t>(stream));
    P2L_TRY_END
}
P2L_EXPORT int p2l_sg2_forward_w(p2l_sg2* m, int b, const float* latent, const float* const* noise, float* img, void* stream) {
    if (!m || b <= 0 || !latent) { set_error("p2l_sg2_forward_w: bad argument"); return -1; }
    P2L_TRY_BEGIN
    return m->g.forward_w(b, latent, noise, img, static_cast<cudaStream_t>(stream));
    P2L_TRY_END
}
P2L_EXPORT int p2l_sg2_backward_w(p2l_sg2* m, int b, const float* dimg, float* dlatent, float* const* dnoise, void* stream) {
    if (!m || b <= 0 || !dimg || !dlatent) { set_error("p2l_sg2_backward_w: bad argument"); return -1; }
    P2L_TRY_BEGIN
    return m->g.backward_w(b, dimg, dlatent, dnoise, static_cast<cudaStream_t>(stream));
    P2L_TRY_END
}
P2L_EXPORT int p2l_sg2_step_w(p2l_sg2* g, p2l_lpips* l, p2l_target* t, int b, const float* latent, const float* const* noise,
                              int want_grad, float grad_scale, const float* dloss, float* loss, float* dlatent,
                              float* const* dnoise, float* img, void* stream) {
    if (!g || !l || !t || b <= 0 || !latent || !loss) { set_error("p2l_sg2_step_w: bad argument"); return -1; }
    if (want_grad && !dlatent) { set_error("p2l_sg2_step_w: want_grad needs dlatent"); return -1; }
    P2L_TRY_BEGIN
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (g->g.forward_w(b, latent, noise, img, st)) return -1;
    const float* im = img ? img : g->g.last_image(b);
    if (l->l.loss_forward(*t->t, b, im, loss, want_grad, st)) return -1;
    if (!want_grad) return 0;
    float* dimg = l->l.unit_grad(*t->t, b);
    if (!dimg) return -1;
    return g->g.backward_w(b, dimg, dlatent, dnoise, st, grad_scale, dloss);
    P2L_TRY_END
}
