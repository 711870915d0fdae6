// Native LPIPS (AlexNet / VGG16 backbones) + weighted L1/L2 pixel term, forward and backward to
// the image. Restates lpips.LPIPS(net, spatial=True) as called from
// /root/reference pix2latent/loss_functions.py:131,142 and the reduction of
// ProjectionLoss / ReconstructionLoss / PerceptualLoss (loss_functions.py:86-148); the CPU
// statement is oracle/lpips.py.
//
// What the reference recomputes every step and this path does once per target:
//   * backbone features of the (constant) target  -> p2l_target, unit-normalised, fp32
//   * five bilinear upsamples to HxW followed by sum(map * W) / sum(W)
//       -> the adjoint of the upsample is applied to W once; the reduction then happens at
//          feature resolution inside the distance kernel (linearity of the upsample).
#include "lpips.h"

#include <cstring>

namespace p2l {

struct LpipsPlan {
    int b = 0, H = 0, W = 0;
    Arena ar;
    struct Lay {
        int Hin, Win, Hout, Wout;        // conv input / output spatial size
        int Hpre, Wpre;                  // size before the optional pool
        act_t *x = nullptr;      // conv input (pooled tensor, im2col buffer or previous F)
        act_t *F = nullptr;      // relu(conv) output
        unsigned char* idx = nullptr;    // pool argmax
        act_t *g = nullptr;      // distance gradient (feature layers)
        act_t *D = nullptr;      // gradient wrt pre-relu conv output
        act_t *dX = nullptr;     // gradient wrt conv input (when a pool or the image follows)
        ConvOp f, d;
    };
    std::vector<Lay> L;
    float* dimg = nullptr;  // unit-upstream d loss_i / d img_i
    float* lossp = nullptr; // [b][nslots] per-block partial sums of the loss (pixel term, then each feature layer)
    int nslots = 0, slot_l1 = 0, slot_feat[8] = {0};
    bool grad_ready = false;
};

Lpips::~Lpips() {}

int Lpips::finalize() {
    if (finalized) return 0;
    convs.clear();
    if (net == P2L_LPIPS_ALEX) {
        // torchvision alexnet.features indices 0,3,6,8,10; slices per lpips/pretrained_networks.py
        const int cfg[5][6] = {{3, 64, 11, 4, 2, 0}, {64, 192, 5, 1, 2, 1}, {192, 384, 3, 1, 1, 1}, {384, 256, 3, 1, 1, 0}, {256, 256, 3, 1, 1, 0}};
        const int idx[5] = {0, 3, 6, 8, 10};
        for (int k = 0; k < 5; ++k) {
            LConv c{};
            c.Cin = cfg[k][0]; c.Cout = cfg[k][1]; c.k = cfg[k][2]; c.stride = cfg[k][3]; c.pad = cfg[k][4];
            c.pool_before = cfg[k][5]; c.pool_k = 3; c.pool_s = 2; c.feat = k;
            c.name = "net.slice" + std::to_string(k + 1) + "." + std::to_string(idx[k]);
            convs.push_back(c);
        }
        nfeat = 5;
    } else if (net == P2L_LPIPS_VGG) {
        const int widths[5] = {64, 128, 256, 512, 512};
        const int counts[5] = {2, 2, 3, 3, 3};
        int i = 0, cin = 3;
        for (int s = 0; s < 5; ++s) {
            if (s > 0) ++i;  // the MaxPool2d module index
            for (int j = 0; j < counts[s]; ++j) {
                LConv c{};
                c.Cin = cin; c.Cout = widths[s]; c.k = 3; c.stride = 1; c.pad = 1;
                c.pool_before = (s > 0 && j == 0); c.pool_k = 2; c.pool_s = 2;
                c.feat = (j == counts[s] - 1) ? s : -1;
                c.name = "net.slice" + std::to_string(s + 1) + "." + std::to_string(i);
                convs.push_back(c);
                cin = widths[s];
                i += 2;
            }
        }
        nfeat = 5;
    } else {
        set_error("lpips: unknown net %d", net);
        return -1;
    }
    for (size_t j = 0; j < convs.size(); ++j) {
        LConv& c = convs[j];
        const auto* w = stage.get(c.name + ".weight", (long)c.Cout * c.Cin * c.k * c.k);
        const auto* bsv = stage.get(c.name + ".bias", c.Cout);
        if (!w || !bsv) return -1;
        c.bias = upload(weights, *bsv);
        if (j == 0 && net == P2L_LPIPS_ALEX) {
            // im2col GEMM: K = 3*11*11 = 363 -> 384, k = (c*11 + r)*11 + s  (torch weight flatten order)
            c.Kp = 384;
            std::vector<act_t> f((size_t)c.Cout * c.Kp, host_f2bf(0.f)), t((size_t)c.Kp * c.Cout, host_f2bf(0.f));
            for (int o = 0; o < c.Cout; ++o)
                for (int k = 0; k < 363; ++k) {
                    f[(size_t)o * c.Kp + k] = host_f2bf((*w)[(size_t)o * 363 + k]);
                    t[(size_t)k * c.Cout + o] = host_f2bf((*w)[(size_t)o * 363 + k]);
                }
            c.w = upload(weights, f);
            c.wt = upload(weights, t);
        } else if (j == 0) {
            // VGG conv1_1: 3 input channels zero-padded to 64
            c.Kp = 64;
            std::vector<float> wp((size_t)c.Cout * 64 * 9, 0.f);
            for (int o = 0; o < c.Cout; ++o)
                for (int ci = 0; ci < 3; ++ci)
                    for (int r = 0; r < 9; ++r) wp[((size_t)o * 64 + ci) * 9 + r] = (*w)[((size_t)o * 3 + ci) * 9 + r];
            c.w = upload(weights, pack_conv_fwd(wp, c.Cout, 64, 3, 3));
            c.wt = upload(weights, pack_conv_dgrad(wp, c.Cout, 64, 3, 3));
        } else {
            c.w = upload(weights, pack_conv_fwd(*w, c.Cout, c.Cin, c.k, c.k));
            c.wt = upload(weights, pack_conv_dgrad(*w, c.Cout, c.Cin, c.k, c.k));
        }
        if (c.feat >= 0) {
            const auto* l = stage.get("lin" + std::to_string(c.feat) + ".weight", c.Cout);
            if (!l) return -1;
            lin[c.feat] = upload(weights, *l);
            chns[c.feat] = c.Cout;
        }
    }
    if (weights.failed) return -1;
    stage.t.clear();
    finalized = true;
    return 0;
}

static int conv_out(int n, int k, int s, int p) { return (n + 2 * p - k) / s + 1; }

int Lpips::feature_dims(int H, int W, int* fh, int* fw) const {
    int h = H, w = W;
    for (const LConv& c : convs) {
        if (c.pool_before) { h = conv_out(h, c.pool_k, c.pool_s, 0); w = conv_out(w, c.pool_k, c.pool_s, 0); }
        h = conv_out(h, c.k, c.stride, c.pad);
        w = conv_out(w, c.k, c.stride, c.pad);
        if (h < 1 || w < 1) { set_error("lpips: image %dx%d too small for the backbone", H, W); return -1; }
        if (c.feat >= 0) { fh[c.feat] = h; fw[c.feat] = w; }
    }
    return 0;
}

static int pick_bn_l(int Cout, long m_tiles, long K = 1 << 20) {
    const int cands[3] = {256, 128, 64};
    for (int k = (K <= 1024 ? 1 : 0); k < 3; ++k)
        if (Cout % cands[k] == 0 && m_tiles * (Cout / cands[k]) >= num_sms()) return cands[k];
    for (int k = 2; k >= 0; --k)
        if (Cout % cands[k] == 0) return cands[k];
    return 64;
}

LpipsPlan* Lpips::plan(int b, int H, int W) {
    const long key = ((long)b << 40) | ((long)H << 20) | W;
    auto it = plans.find(key);
    if (it != plans.end()) return it->second.get();
    std::shared_ptr<LpipsPlan> pp(new LpipsPlan());
    LpipsPlan& P = *pp;
    P.b = b; P.H = H; P.W = W;
    Arena& ar = P.ar;
    typedef act_t bf;
    const int n = (int)convs.size();
    P.L.resize(n);
    P.dimg = ar.alloc<float>((size_t)b * 3 * H * W);
    int h = H, w = W;
    for (int j = 0; j < n; ++j) {
        const LConv& c = convs[j];
        LpipsPlan::Lay& l = P.L[j];
        l.Hpre = h; l.Wpre = w;
        if (c.pool_before) {
            h = conv_out(h, c.pool_k, c.pool_s, 0);
            w = conv_out(w, c.pool_k, c.pool_s, 0);
            l.x = ar.alloc<bf>((size_t)b * h * w * c.Cin);
            l.idx = ar.alloc<unsigned char>((size_t)b * h * w * c.Cin);
            l.dX = ar.alloc<bf>((size_t)b * h * w * c.Cin);
        }
        l.Hin = h; l.Win = w;
        h = conv_out(h, c.k, c.stride, c.pad);
        w = conv_out(w, c.k, c.stride, c.pad);
        if (h < 1 || w < 1) { set_error("lpips: image too small"); return nullptr; }
        l.Hout = h; l.Wout = w;
        if (j == 0) {
            if (net == P2L_LPIPS_ALEX) {
                l.x = ar.alloc<bf>((size_t)b * h * w * c.Kp);   // im2col
                l.dX = ar.alloc<bf>((size_t)b * h * w * c.Kp);
            } else {
                l.x = ar.alloc<bf>((size_t)b * H * W * c.Kp);   // padded NHWC image
                l.dX = ar.alloc<bf>((size_t)b * H * W * c.Kp);
            }
        } else if (!c.pool_before) {
            l.x = P.L[j - 1].F;
        }
        l.F = ar.alloc<bf>((size_t)b * h * w * c.Cout);
        l.D = ar.alloc<bf>((size_t)b * h * w * c.Cout);
        if (c.feat >= 0) l.g = ar.alloc<bf>((size_t)b * h * w * c.Cout);
    }
    // the last conv's pre-relu gradient IS its (relu-masked) distance gradient
    P.L[n - 1].D = P.L[n - 1].g;
    // loss partial slots: every block of the pixel-term / distance kernels owns one (no atomics: reproducible sums)
    P.slot_l1 = 0;
    P.nslots = k_l1_loss_slots(3 * H * W);
    for (int j = 0; j < n; ++j) {
        if (convs[j].feat < 0) continue;
        P.slot_feat[convs[j].feat] = P.nslots;
        P.nslots += k_lpips_dist_slots(P.L[j].Hout * P.L[j].Wout);
    }
    P.lossp = ar.alloc<float>((size_t)b * P.nslots);
    if (ar.failed) return nullptr;
    auto m_tiles = [&](int hh, int ww) {
        int tw = 1; while (tw < ww) tw <<= 1; if (tw > 16) tw = 16;
        int th = 1; while (th < hh) th <<= 1; if (th > 128 / tw) th = 128 / tw;
        const int nb = 128 / (tw * th);
        return (long)((ww + tw - 1) / tw) * ((hh + th - 1) / th) * ((b + nb - 1) / nb);
    };
    for (int j = 0; j < n; ++j) {
        const LConv& c = convs[j];
        LpipsPlan::Lay& l = P.L[j];
        {   // forward: F = relu(conv(x) + bias)
            ConvDesc d;
            if (j == 0 && net == P2L_LPIPS_ALEX) {
                d.A = l.x; d.A_N = b; d.A_H = l.Hout; d.A_W = l.Wout; d.A_C = c.Kp; d.Cin = c.Kp;
                d.kh = d.kw = 1; d.pad_h = d.pad_w = 0;
            } else {
                const int cin = (j == 0) ? c.Kp : c.Cin;
                d.A = l.x; d.A_N = b; d.A_H = l.Hin; d.A_W = l.Win; d.A_C = cin; d.Cin = cin;
                d.kh = d.kw = c.k; d.pad_h = d.pad_w = c.pad;
            }
            d.B = c.w; d.Cout = c.Cout;
            d.NI = b; d.H = l.Hout; d.W = l.Wout; d.mode = EPI_FWD;
            d.BN = pick_bn_l(c.Cout, m_tiles(l.Hout, l.Wout));
            d.epi.bias = c.bias; d.epi.relu = 1; d.epi.act = l.F; d.epi.act_C = c.Cout;
            if (conv_op_build(&l.f, d)) return nullptr;
            l.f.p.tile_reverse = (get_option("serpentine") != 0) ? ((j + 1) & 1) : 0;  // the rgb head before conv 0 walked backwards
        }
        {   // backward: gradient wrt the conv input
            ConvDesc d;
            int cin_eff;
            d.A = l.D; d.A_N = b; d.A_H = l.Hout; d.A_W = l.Wout; d.A_C = c.Cout; d.Cin = c.Cout;
            d.B = c.wt;
            if (j == 0 && net == P2L_LPIPS_ALEX) {
                cin_eff = c.Kp; d.kh = d.kw = 1; d.pad_h = d.pad_w = 0;
                d.NI = b; d.H = l.Hout; d.W = l.Wout;
            } else {
                cin_eff = (j == 0) ? c.Kp : c.Cin; d.kh = d.kw = c.k; d.pad_h = d.pad_w = c.pad;
                d.NI = b; d.H = l.Hin; d.W = l.Win;
            }
            d.Cout = cin_eff;
            d.mode = EPI_BWD;
            d.BN = pick_bn_l(cin_eff, m_tiles(d.H, d.W));
            if (j == 0 || c.pool_before) {
                d.epi.dx = l.dX; d.epi.dx_C = cin_eff;
            } else {
                // straight into the previous layer's pre-relu gradient: relu mask + its distance gradient
                d.epi.saved = P.L[j - 1].F; d.epi.saved_C = cin_eff;
                if (convs[j - 1].feat >= 0) { d.epi.addin = P.L[j - 1].g; d.epi.addin_C = cin_eff; d.epi.addin_climit = cin_eff; }
                d.epi.dx = P.L[j - 1].D; d.epi.dx_C = cin_eff;
            }
            if (conv_op_build(&l.d, d)) return nullptr;
            l.d.p.tile_reverse = (get_option("serpentine") != 0) ? ((n - 1 - j) & 1) : 0;
        }
    }
    LpipsPlan* raw = pp.get();
    plans[key] = pp;
    return raw;
}

int Lpips::features(LpipsPlan& P, const float* img, cudaStream_t st) {
    const int n = (int)convs.size();
    const int b = P.b;
    for (int j = 0; j < n; ++j) {
        const LConv& c = convs[j];
        LpipsPlan::Lay& l = P.L[j];
        if (j == 0) {
            if (net == P2L_LPIPS_ALEX) k_im2col_alex1(img, l.x, b, P.H, P.W, l.Hout, l.Wout, c.Kp, st);
            else k_img_to_nhwc_scaled(img, l.x, b, P.H, P.W, c.Kp, st);
        } else if (c.pool_before) {
            k_maxpool_fwd(P.L[j - 1].F, l.x, l.idx, b, l.Hpre, l.Wpre, c.Cin, l.Hin, l.Win, c.pool_k, c.pool_s, st);
        }
        if (conv_op_launch(l.f, st)) return -1;
    }
    return 0;
}

// ----------------------------------------------------------------------------- target
Target* Lpips::make_target(const float* target, const float* weight, const float* mask, int H, int W, int rec_type,
                           float rec_weight, float per_weight, cudaStream_t st) {
    if (!finalized) { set_error("lpips: target before finalize"); return nullptr; }
    std::unique_ptr<Target> tp(new Target());
    Target& T = *tp;
    T.m = this; T.H = H; T.W = W; T.rec_type = rec_type; T.rec_w = rec_weight; T.per_w = per_weight;
    const size_t n3 = (size_t)3 * H * W;
    T.target = T.ar.alloc<float>(n3);
    cudaMemcpyAsync(T.target, target, n3 * 4, cudaMemcpyDefault, st);
    if (weight) { T.weight = T.ar.alloc<float>(n3); cudaMemcpyAsync(T.weight, weight, n3 * 4, cudaMemcpyDefault, st); }
    if (mask) { T.mask = T.ar.alloc<float>(n3); cudaMemcpyAsync(T.mask, mask, n3 * 4, cudaMemcpyDefault, st); }
    T.wsum = T.ar.alloc<float>((size_t)H * W);
    T.total = T.ar.alloc<float>(4);
    k_weight_sum(T.weight, T.mask, T.wsum, T.total, H * W, st);
    float total = 0.f;
    if (cudaMemcpyAsync(&total, T.total, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
        set_error("target: reading sum(W) failed: %s", cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    if (!(total > 0.f)) { set_error("target: sum of weights is %g", total); return nullptr; }
    T.sumW = total;
    // scalars used by the kernels: total_rec = sumW / rec_weight
    const float scal[2] = {rec_weight != 0.f ? total / rec_weight : 1.f, 0.f};
    cudaMemcpyAsync(T.total + 1, scal, 8, cudaMemcpyHostToDevice, st);
    if (feature_dims(H, W, T.fh, T.fw)) return nullptr;
    // target features through the same kernels (b = 1)
    LpipsPlan* P = plan(1, H, W);
    if (!P) return nullptr;
    if (features(*P, T.target, st)) return nullptr;
    for (size_t j = 0; j < convs.size(); ++j) {
        const int f = convs[j].feat;
        if (f < 0) continue;
        const int hw = T.fh[f] * T.fw[f];
        T.tfeat[f] = T.ar.alloc<float>((size_t)hw * chns[f]);
        T.wadj[f] = T.ar.alloc<float>(hw);
        k_lpips_normalize(P->L[j].F, T.tfeat[f], hw, chns[f], st);
        k_upsample_adjoint(T.wsum, T.wadj[f], H, W, T.fh[f], T.fw[f], per_weight / total, st);
    }
    if (T.ar.failed) return nullptr;
    if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("target: %s", cudaGetErrorString(cudaGetLastError())); return nullptr; }
    return tp.release();
}

// ----------------------------------------------------------------------------- loss
int Lpips::loss_forward(Target& T, int b, const float* img, float* loss, int want_grad, cudaStream_t st) {
    std::vector<Target*> Ts((size_t)b, &T);
    return loss_forward_multi(Ts.data(), b, img, loss, want_grad, st);
}

// One target PER CANDIDATE (transform search, pix2latent/transform/transform_optimizer.py:165-255: every
// candidate's target / weight is its own affine resample): the generator-side work (features, dgrad) is
// batched as before; only the target-dependent kernels (pixel term, layer distances) run once per run of
// consecutive identical targets. All targets must share the resolution and the loss configuration.
int Lpips::loss_forward_multi(Target* const* Ts, int b, const float* img, float* loss, int want_grad, cudaStream_t st) {
    Target& T = *Ts[0];
    for (int i = 1; i < b; ++i) {
        const Target& U = *Ts[i];
        if (U.H != T.H || U.W != T.W || U.rec_type != T.rec_type || U.rec_w != T.rec_w || U.per_w != T.per_w || U.m != T.m) {
            set_error("lpips: per-candidate targets must share resolution and loss configuration");
            return -1;
        }
    }
    LpipsPlan* Pp = plan(b, T.H, T.W);
    if (!Pp) return -1;
    LpipsPlan& P = *Pp;
    const int n = (int)convs.size();
    const int HW = T.H * T.W;
    // runs of consecutive identical targets: [r0[k], r0[k+1])
    std::vector<int> r0;
    for (int i = 0; i < b; ++i)
        if (i == 0 || Ts[i] != Ts[i - 1]) r0.push_back(i);
    r0.push_back(b);
    const int nr = (int)r0.size() - 1;
    P2L_CUDA_CHECK(cudaMemsetAsync(P.lossp, 0, (size_t)b * P.nslots * sizeof(float), st));
    // pixel term (also initialises dimg)
    if (T.rec_w != 0.f) {
        for (int k = 0; k < nr; ++k) {
            const Target& U = *Ts[r0[k]];
            const int i0 = r0[k], nb = r0[k + 1] - r0[k];
            k_l1_loss(img + (size_t)i0 * 3 * HW, U.target, U.weight, U.mask, U.total + 1, P.lossp + (size_t)i0 * P.nslots + P.slot_l1,
                      P.nslots, want_grad ? P.dimg + (size_t)i0 * 3 * HW : nullptr, nb, 3 * HW, HW, U.rec_type == 2, st);
        }
    } else if (want_grad) {
        P2L_CUDA_CHECK(cudaMemsetAsync(P.dimg, 0, (size_t)b * 3 * HW * sizeof(float), st));
    }
    P.grad_ready = false;
    if (T.per_w == 0.f) {
        k_loss_reduce(P.lossp, P.nslots, P.nslots, loss, b, st);
        P.grad_ready = want_grad;
        return 0;
    }
    if (features(P, img, st)) return -1;
    for (int j = 0; j < n; ++j) {
        const int f = convs[j].feat;
        if (f < 0) continue;
        const size_t per = (size_t)T.fh[f] * T.fw[f] * chns[f];
        for (int k = 0; k < nr; ++k) {
            const Target& U = *Ts[r0[k]];
            const int i0 = r0[k], nb = r0[k + 1] - r0[k];
            k_lpips_dist(P.L[j].F + i0 * per, U.tfeat[f], lin[f], U.wadj[f], P.lossp + (size_t)i0 * P.nslots + P.slot_feat[f], P.nslots,
                         want_grad ? P.L[j].g + i0 * per : nullptr, nb, T.fh[f] * T.fw[f], chns[f], grad_scale(), st);
        }
    }
    k_loss_reduce(P.lossp, P.nslots, P.nslots, loss, b, st);
    if (!want_grad) return 0;
    // ---- backward through the backbone (dgrad only)
    // last conv's pre-relu gradient is its (already relu-masked) distance gradient
    for (int j = n - 1; j >= 0; --j) {
        const LConv& c = convs[j];
        LpipsPlan::Lay& l = P.L[j];
        if (conv_op_launch(l.d, st)) return -1;
        if (j == 0) {
            if (net == P2L_LPIPS_ALEX) k_col2im_alex1(l.dX, P.dimg, b, T.H, T.W, l.Hout, l.Wout, c.Kp, 1, 1.f / grad_scale(), st);
            else k_nhwc_to_dimg_scaled(l.dX, c.Kp, P.dimg, b, T.H, T.W, 1, 1.f / grad_scale(), st);
        } else if (c.pool_before) {
            const LpipsPlan::Lay& pl = P.L[j - 1];
            k_maxpool_bwd(l.dX, l.idx, pl.F, convs[j - 1].feat >= 0 ? pl.g : nullptr, pl.D, b, l.Hpre, l.Wpre, c.Cin, l.Hin,
                          l.Win, c.pool_k, c.pool_s, st);
        }
    }
    P.grad_ready = true;
    return 0;
}

int Lpips::loss_backward(Target& T, int b, const float* dloss, float* dimg, cudaStream_t st) {
    LpipsPlan* Pp = plan(b, T.H, T.W);
    if (!Pp || !Pp->grad_ready) { set_error("lpips: loss_backward without loss_forward(want_grad=1)"); return -1; }
    const size_t n = (size_t)3 * T.H * T.W;
    if (dimg != Pp->dimg) P2L_CUDA_CHECK(cudaMemcpyAsync(dimg, Pp->dimg, (size_t)b * n * 4, cudaMemcpyDeviceToDevice, st));
    k_scale_rows(dimg, dloss, b, (long)n, st);
    return 0;
}

float* Lpips::unit_grad(Target& T, int b) {
    LpipsPlan* Pp = plan(b, T.H, T.W);
    return Pp ? Pp->dimg : nullptr;
}

double Lpips::flops(int b, int H, int W, int backward) {
    LpipsPlan* P = plan(b, H, W);
    if (!P) return 0;
    double f = 0;
    for (auto& l : P->L) f += backward ? l.d.flops : l.f.flops;
    return f;
}

}  // namespace p2l
