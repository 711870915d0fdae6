// Native StyleGAN2 generator (see sg2.h). Layer algebra follows oracle/stylegan2.py, which restates
// rosinality/stylegan2-pytorch model.py as reached via pix2latent/model/stylegan2.py:116-119.
//
// Execution map (reference op -> here; sg_epilogue.cuh has the arithmetic):
//   ModulatedConv2d (grouped conv with b*Cout per-sample filters) + NoiseInjection + FusedLeakyReLU
//        -> ONE tcgen05 convolution with SHARED weights per layer: the per-sample modulation is applied by the PREVIOUS
//           layer's epilogue (it writes x_l and A_{l+1} = x_l * s_{l+1}), demodulation + noise + bias + leaky-ReLU are
//           this layer's epilogue
//   conv_transpose2d(stride 2) + Blur (upfirdn2d)
//        -> the two linear maps composed into ONE 3x3 filter per output phase (py, px): a plain 3x3 convolution on the
//           LOW-resolution grid with 4*Cout columns whose stores go depth-to-space; its gradient arrives
//           space-to-depth, so the dgrad is a plain 3x3 convolution too. No zero-inserted grid, no FIR pass.
//   backward: layer l's dgrad epilogue = modulation-backward of layer l + activation / noise / demodulation-backward of
//           layer l-1 (two per-(sample, channel) sums in the deterministic partial slots)
//   ToRGB (1x1 modulated conv, no demod) + Upsample skip (upfirdn2d CUDA op) -> k_sg_torgb_fwd / _bwd
#include "sg2.h"

#include <cmath>
#include <cstring>

namespace p2l {

struct SG2Plan {
    int b = 0;
    Arena ar;
    float *z = nullptr, *h[9] = {nullptr}, *g0 = nullptr, *g1 = nullptr;
    float *s_all = nullptr, *ds_all = nullptr, *dm_all = nullptr, *ddm_all = nullptr, *dw = nullptr;
    struct Lay {
        act_t* x = nullptr;      // x_l: the layer's output (saved for the backward pass and read by ToRGB)
        float* sp = nullptr;     // partial slots of (ds_l, ddm_{l-1}) filled by this layer's dgrad epilogue
        int sp_parts = 0;
        ConvOp f, d;
    };
    std::vector<Lay> L;
    act_t* A[2] = {nullptr, nullptr};   // modulated conv inputs A_l = x_{l-1} * s_l (ping-pong: A_l is read by layer l only)
    act_t* G[2] = {nullptr, nullptr};   // dgrad inputs G_l (ping-pong)
    act_t *dxrgb = nullptr, *dx2 = nullptr, *dA0 = nullptr;
    StatSeg* segs = nullptr;
    int nsegs = 0;
    void *demod_tab = nullptr, *demod_bwd_tab = nullptr;   // batched per-layer demodulation GEMMs (b <= 24)
    int demod_maxJ = 0, demod_bwd_maxJ = 0;
    std::vector<float*> rgb, weff, dweff, drgb;
    float* img = nullptr;
    float* scratch = nullptr;  // block partials of the per-(sample, channel) reductions (sg2_kernels.h)
    std::vector<const float*> noise_ptrs;   // the last forward's noise images (read again by the backward epilogues)
    bool forward_done = false;
    int mode = 0;  // 0: z search (mapping network ran), 1: w / w+ search (styles from the latent rows)
};

SG2::~SG2() {}

static int pick_bn_sg(int Cout, long m_tiles, long K) {
    const int cands[3] = {256, 128, 64};
    for (int k = (K <= 1024 ? 1 : 0); k < 3; ++k)
        if (Cout % cands[k] == 0 && m_tiles * (Cout / cands[k]) >= num_sms()) return cands[k];
    for (int k = 2; k >= 0; --k)
        if (Cout % cands[k] == 0) return cands[k];
    return 64;
}

int SG2::finalize() {
    if (finalized) return 0;
    sdim = cfg.style_dim;
    n_mlp = cfg.n_mlp;
    log_size = 0;
    while ((1 << log_size) < cfg.size) ++log_size;
    if ((1 << log_size) != cfg.size || log_size < 3 || log_size > 10 || n_mlp != 8) { set_error("sg2: unsupported size %d / n_mlp %d", cfg.size, n_mlp); return -1; }
    num_layers = (log_size - 2) * 2 + 1;
    // ---- mapping network: EqualLinear(512, 512, lr_mul=0.01, fused_lrelu)
    const float lr = 0.01f;
    map_scale = (1.f / std::sqrt((float)sdim)) * lr;
    for (int k = 1; k <= n_mlp; ++k) {
        const auto* w = stage.get("style." + std::to_string(k) + ".weight", (long)sdim * sdim);
        const auto* bsv = stage.get("style." + std::to_string(k) + ".bias", sdim);
        if (!w || !bsv) return -1;
        std::vector<float> t((size_t)sdim * sdim), bb(sdim);
        for (int o = 0; o < sdim; ++o)
            for (int i = 0; i < sdim; ++i) t[(size_t)i * sdim + o] = (*w)[(size_t)o * sdim + i];
        for (int o = 0; o < sdim; ++o) bb[o] = (*bsv)[o] * lr;
        map_WT.push_back(upload(weights, t));
        map_W.push_back(upload(weights, *w));
        map_b.push_back(upload(weights, bb));
    }
    // ---- layer table
    auto ch = [&](int res_log) { return cfg.channels[res_log - 2]; };
    convs.clear();
    rgbs.clear();
    int s_off = 0, dm_off = 0;
    std::vector<std::string> conv_names, rgb_names;
    {
        Conv c{}; c.Cin = ch(2); c.Cout = ch(2); c.Hin = 4; c.Hout = 4; c.up = 0;
        convs.push_back(c); conv_names.push_back("conv1");
        Rgb r{}; r.Cin = ch(2); r.H = 4; rgbs.push_back(r); rgb_names.push_back("to_rgb1");
    }
    int in_c = ch(2);
    for (int i = 3; i <= log_size; ++i) {
        const int out_c = ch(i), res = 1 << i;
        Conv a{}; a.Cin = in_c; a.Cout = out_c; a.Hin = res / 2; a.Hout = res; a.up = 1;
        Conv bq{}; bq.Cin = out_c; bq.Cout = out_c; bq.Hin = res; bq.Hout = res; bq.up = 0;
        convs.push_back(a); conv_names.push_back("convs." + std::to_string(2 * (i - 3)));
        convs.push_back(bq); conv_names.push_back("convs." + std::to_string(2 * (i - 3) + 1));
        Rgb r{}; r.Cin = out_c; r.H = res; rgbs.push_back(r); rgb_names.push_back("to_rgbs." + std::to_string(i - 3));
        in_c = out_c;
    }
    for (auto& c : convs) {
        if (c.Cin % 64 || c.Cout % 64) { set_error("sg2: channel counts must be multiples of 64 (got %d -> %d)", c.Cin, c.Cout); return -1; }
        c.s_off = s_off; s_off += c.Cin;
        c.dm_off = dm_off; dm_off += c.Cout;
    }
    for (auto& r : rgbs) { r.s_off = s_off; s_off += r.Cin; }
    S = s_off;
    DM = dm_off;
    // ---- modulation affines (EqualLinear(512, Cin, bias_init=1), lr_mul 1)
    std::vector<float> h_aff((size_t)S * sdim), h_affT((size_t)S * sdim), h_affb(S);
    auto put_aff = [&](const std::string& pre, int Cin, int off) -> int {
        const auto* w = stage.get(pre + ".conv.modulation.weight", (long)Cin * sdim);
        const auto* bsv = stage.get(pre + ".conv.modulation.bias", Cin);
        if (!w || !bsv) return -1;
        for (int i = 0; i < Cin; ++i) {
            h_affb[off + i] = (*bsv)[i];
            for (int k = 0; k < sdim; ++k) {
                h_aff[(size_t)(off + i) * sdim + k] = (*w)[(size_t)i * sdim + k];
                h_affT[(size_t)k * S + off + i] = (*w)[(size_t)i * sdim + k];
            }
        }
        return 0;
    };
    for (size_t l = 0; l < convs.size(); ++l) {
        Conv& c = convs[l];
        const std::string& pre = conv_names[l];
        if (put_aff(pre, c.Cin, c.s_off)) return -1;
        const auto* w = stage.get(pre + ".conv.weight", (long)c.Cout * c.Cin * 9);
        const auto* nw = stage.get(pre + ".noise.weight", 1);
        const auto* bsv = stage.get(pre + ".activate.bias", c.Cout);
        if (!w || !nw || !bsv) return -1;
        const float scale = 1.f / std::sqrt((float)c.Cin * 9.f);
        std::vector<float> wsq((size_t)c.Cout * c.Cin), wsqT((size_t)c.Cout * c.Cin);
        for (int o = 0; o < c.Cout; ++o)
            for (int i = 0; i < c.Cin; ++i) {
                double q = 0;
                for (int t = 0; t < 9; ++t) {
                    const double v = (double)(*w)[((size_t)o * c.Cin + i) * 9 + t] * scale;
                    q += v * v;   // demodulation uses the 3x3 weights themselves (before the transposed conv / blur)
                }
                wsq[(size_t)o * c.Cin + i] = (float)q;
                wsqT[(size_t)i * c.Cout + o] = (float)q;
            }
        if (!c.up) {
            std::vector<float> ws(w->size());
            for (size_t t = 0; t < w->size(); ++t) ws[t] = (*w)[t] * scale;
            c.w = upload(weights, pack_conv_fwd(ws, c.Cout, c.Cin, 3, 3));
            c.wt = upload(weights, pack_conv_dgrad(ws, c.Cout, c.Cin, 3, 3));
        } else {
            // conv_transpose2d(stride 2): D'[2i + ky] += W[ky] in[i]; Blur (upfirdn2d, 4-tap FIR kf, pad (1, 1)):
            // out[Y] = sum_u kf[u] D'[Y + u - 1]. Composed, per axis and output phase py = Y & 1 (Y = 2i + py):
            //   out[2i + py] = sum_{d in -1..1} Wc_py[d] in[i + d],   Wc_py[d] = sum_{u, ky : py + u - 1 - ky = 2d} kf[u] W[ky]
            // i.e. four plain 3x3 filters on the low-resolution grid, one per phase (py, px): a conv with 4*Cout outputs.
            const double kf[4] = {0.25, 0.75, 0.75, 0.25};   // [1,3,3,1] / 8 * 2 per axis (upsample_factor 2)
            std::vector<float> wc((size_t)4 * c.Cout * c.Cin * 9, 0.f);
            for (int py = 0; py < 2; ++py)
                for (int px = 0; px < 2; ++px)
                    for (int u = 0; u < 4; ++u)
                        for (int ky = 0; ky < 3; ++ky) {
                            const int ey = py + u - 1 - ky;
                            if (ey & 1) continue;
                            const int dy = ey / 2;   // -1, 0, 1
                            for (int v = 0; v < 4; ++v)
                                for (int kx = 0; kx < 3; ++kx) {
                                    const int ex = px + v - 1 - kx;
                                    if (ex & 1) continue;
                                    const int dx = ex / 2;
                                    const double kk = kf[u] * kf[v] * scale;
                                    for (int o = 0; o < c.Cout; ++o)
                                        for (int i = 0; i < c.Cin; ++i)
                                            wc[((((size_t)(py * 2 + px) * c.Cout + o) * c.Cin + i) * 3 + (dy + 1)) * 3 + (dx + 1)] +=
                                                (float)(kk * (*w)[(((size_t)o * c.Cin + i) * 3 + ky) * 3 + kx]);
                                }
                        }
            c.w = upload(weights, pack_conv_fwd(wc, 4 * c.Cout, c.Cin, 3, 3));
            c.wt = upload(weights, pack_conv_dgrad(wc, 4 * c.Cout, c.Cin, 3, 3));
        }
        c.wsq = upload(weights, wsq);
        c.wsqT = upload(weights, wsqT);
        c.noise_w = upload(weights, *nw);
        c.bias = upload(weights, *bsv);
    }
    for (size_t t = 0; t < rgbs.size(); ++t) {
        Rgb& r = rgbs[t];
        const std::string& pre = rgb_names[t];
        if (put_aff(pre, r.Cin, r.s_off)) return -1;
        const auto* w = stage.get(pre + ".conv.weight", (long)3 * r.Cin);
        const auto* bsv = stage.get(pre + ".bias", 3);
        if (!w || !bsv) return -1;
        r.Wr = upload(weights, *w);
        r.bias = upload(weights, *bsv);
        r.scale = 1.f / std::sqrt((float)r.Cin);
    }
    aff = upload(weights, h_aff);
    affT = upload(weights, h_affT);
    aff_b = upload(weights, h_affb);
    {   // constant input [1, C0, 4, 4] -> NHWC bf16
        const int C0 = convs[0].Cin;
        const auto* w = stage.get("input.input", (long)C0 * 16);
        if (!w) return -1;
        std::vector<act_t> t((size_t)16 * C0);
        for (int c = 0; c < C0; ++c)
            for (int p = 0; p < 16; ++p) t[(size_t)p * C0 + c] = host_f2bf((*w)[(size_t)c * 16 + p]);
        const_in = upload(weights, t);
    }
    if (weights.failed) return -1;
    stage.t.clear();
    finalized = true;
    return 0;
}

static bool has_rgb(int l) { return l == 0 || (l % 2 == 0); }   // ToRGB t reads x_0 (t = 0) or x_{2t}

SG2Plan* SG2::plan(int b) {
    auto it = plans.find(b);
    if (it != plans.end()) return it->second.get();
    std::shared_ptr<SG2Plan> pp(new SG2Plan());
    SG2Plan& P = *pp;
    P.b = b;
    Arena& ar = P.ar;
    typedef act_t bf;
    P.z = ar.alloc<float>((size_t)b * sdim);
    for (int k = 0; k <= n_mlp; ++k) P.h[k] = ar.alloc<float>((size_t)b * sdim);
    P.g0 = ar.alloc<float>((size_t)b * sdim);
    P.g1 = ar.alloc<float>((size_t)b * sdim);
    P.s_all = ar.alloc<float>((size_t)b * S);
    P.ds_all = ar.alloc<float>((size_t)b * S);
    P.dm_all = ar.alloc<float>((size_t)b * DM);
    P.ddm_all = ar.alloc<float>((size_t)b * DM);
    P.dw = ar.alloc<float>((size_t)b * sdim);
    const int nL = (int)convs.size();
    P.L.resize(nL);
    size_t max_x = 0, max_a = 0;
    long ms = 0;
    for (int l = 0; l < nL; ++l) {
        const Conv& c = convs[l];
        SG2Plan::Lay& q = P.L[l];
        q.x = ar.alloc<bf>((size_t)b * c.Hout * c.Hout * c.Cout);
        max_x = std::max(max_x, (size_t)b * c.Hout * c.Hout * c.Cout);
        max_a = std::max(max_a, (size_t)b * c.Hin * c.Hin * c.Cin);
        if (l > 0) {
            q.sp_parts = conv_stat_parts_max(c.Hin, c.Hin);
            q.sp = ar.alloc<float>((size_t)b * q.sp_parts * 2 * c.Cin);
        }
        ms = std::max(ms, std::max(k_sg_scratch_floats(b, c.Hout, c.Hout, c.Cout), k_sg_scratch_floats(b, c.Hin, c.Hin, c.Cin)));
    }
    for (int k = 0; k < 2; ++k) {
        P.A[k] = ar.alloc<bf>(std::max(max_a, max_x));
        P.G[k] = ar.alloc<bf>(max_x);
    }
    P.dxrgb = ar.alloc<bf>(max_x);
    P.dx2 = ar.alloc<bf>(max_x);
    P.dA0 = ar.alloc<bf>((size_t)b * convs[0].Hin * convs[0].Hin * convs[0].Cin);
    ms = std::max(ms, 8L * 24 * sdim);   // split-K partials of the style-gradient GEMM (k_fc_bwd)
    P.scratch = ar.alloc<float>((size_t)ms);
    const int R = cfg.size;
    for (size_t t = 0; t < rgbs.size(); ++t) {
        P.rgb.push_back(ar.alloc<float>((size_t)b * 3 * rgbs[t].H * rgbs[t].H));
        P.drgb.push_back(ar.alloc<float>((size_t)b * 3 * rgbs[t].H * rgbs[t].H));
        P.weff.push_back(ar.alloc<float>((size_t)b * 3 * rgbs[t].Cin));
        P.dweff.push_back(ar.alloc<float>((size_t)b * 3 * rgbs[t].Cin));
    }
    P.img = ar.alloc<float>((size_t)b * 3 * R * R);
    if (ar.failed) return nullptr;
    auto m_tiles = [&](int hh) {
        int tw = 1; while (tw < hh) tw <<= 1; if (tw > 16) tw = 16;
        int th = 1; while (th < hh) th <<= 1; if (th > 128 / tw) th = 128 / tw;
        const int nb = 128 / (tw * th);
        return (long)((hh + tw - 1) / tw) * ((hh + th - 1) / th) * ((b + nb - 1) / nb);
    };
    std::vector<StatSeg> segs;
    const int serp = get_option("serpentine") != 0 ? 1 : 0;
    for (int l = 0; l < nL; ++l) {
        const Conv& c = convs[l];
        SG2Plan::Lay& q = P.L[l];
        const int Nf = c.up ? 4 * c.Cout : c.Cout;   // GEMM columns: (phase, channel) for up-sampling layers
        {   // forward: x_l = lrelu(dm * conv(A_l) + noise + bias) * sqrt2 ; A_{l+1} = s_{l+1} * x_l
            ConvDesc d;
            d.A = P.A[l & 1]; d.A_N = b; d.A_H = c.Hin; d.A_W = c.Hin; d.A_C = c.Cin; d.Cin = c.Cin;
            d.B = c.w; d.Cout = Nf; d.kh = d.kw = 3; d.pad_h = d.pad_w = 1;
            d.NI = b; d.H = c.Hin; d.W = c.Hin; d.mode = EPI_FWD;
            d.BN = pick_bn_sg(Nf, m_tiles(c.Hin), 9L * c.Cin);
            d.sg = 1;
            ConvGemmParams& e = d.epi;
            e.bias = c.bias;
            e.sg_dm = P.dm_all + c.dm_off; e.sg_ld = DM;
            e.sg_nw = c.noise_w;                      // e.sg_noise: the caller's noise image, set per call
            e.d2s_C = c.up ? c.Cout : 0;
            e.raw = q.x; e.raw_C = c.Cout;
            if (l + 1 < nL) {
                e.aff_a = P.s_all + convs[l + 1].s_off; e.aff_stride = S;
                e.act = P.A[(l + 1) & 1]; e.act_C = c.Cout;
            }
            e.tile_reverse = serp & l;
            if (conv_op_build(&q.f, d)) return nullptr;
        }
        if (l == 0) {
            // layer 0 reads the learned constant: only ds_0 = sum dA * const is needed (k_sg_modulate_bwd)
            ConvDesc d;
            d.A = P.G[l & 1]; d.A_N = b; d.A_H = c.Hout; d.A_W = c.Hout; d.A_C = c.Cout; d.Cin = c.Cout;
            d.B = c.wt; d.Cout = c.Cin; d.kh = d.kw = 3; d.pad_h = d.pad_w = 1;
            d.NI = b; d.H = c.Hin; d.W = c.Hin; d.mode = EPI_BWD;
            d.BN = pick_bn_sg(c.Cin, m_tiles(c.Hin), 9L * c.Cout);
            d.epi.dx = P.dA0; d.epi.dx_C = c.Cin;
            if (conv_op_build(&q.d, d)) return nullptr;
            continue;
        }
        {   // backward: dgrad of layer l; its epilogue also takes the gradient through layer l-1's activation / noise /
            // demodulation and writes G_{l-1} (space-to-depth when layer l-1 up-samples)
            const Conv& pv = convs[l - 1];
            ConvDesc d;
            d.A = P.G[l & 1]; d.A_N = b; d.A_H = c.Hin; d.A_W = c.Hin; d.A_C = Nf; d.Cin = Nf;   // up: G_l arrives space-to-depth
            d.B = c.wt; d.Cout = c.Cin; d.kh = d.kw = 3; d.pad_h = d.pad_w = 1;
            d.NI = b; d.H = c.Hin; d.W = c.Hin; d.mode = EPI_BWD;
            d.BN = pick_bn_sg(c.Cin, m_tiles(c.Hin), 9L * Nf);
            d.sg = 1;
            ConvGemmParams& e = d.epi;
            e.saved = P.L[l - 1].x; e.saved_C = c.Cin;
            e.aff_a = P.s_all + c.s_off; e.aff_stride = S;
            e.sg_dm = P.dm_all + pv.dm_off; e.sg_ld = DM;
            e.sg_bias = pv.bias; e.sg_nw = pv.noise_w;   // e.sg_noise (layer l-1's noise), e.addin, e.dx2: set per call
            e.statp = q.sp; e.statp_parts = q.sp_parts; e.statp_C = c.Cin;
            e.dx = P.G[(l - 1) & 1]; e.dx_C = c.Cin; e.s2d = pv.up ? 1 : 0;
            e.addin_C = c.Cin; e.addin_climit = c.Cin;
            e.tile_reverse = serp & (l ^ 1) & 1;
            if (conv_op_build(&q.d, d)) return nullptr;
            for (int c0 = 0; c0 < c.Cin; c0 += 32) segs.push_back({q.sp, q.d.stat_parts, q.sp_parts, c.Cin, c.s_off, c0, pv.dm_off});
        }
    }
    P.nsegs = (int)segs.size();
    P.segs = upload(ar, segs);
    if (b <= 24) {
        std::vector<unsigned char> h0(k_sg_batch_bytes(nL)), h1(k_sg_batch_bytes(nL));
        for (int l = 0; l < nL; ++l) {
            const Conv& c = convs[l];
            k_sg_demod_desc(h0.data(), l, P.s_all + c.s_off, S, c.wsqT, P.dm_all + c.dm_off, DM, b, c.Cin, c.Cout);
            k_sg_demod_bwd_desc(h1.data(), l, P.ddm_all + c.dm_off, P.dm_all + c.dm_off, DM, P.s_all + c.s_off, S, c.wsq,
                                P.ds_all + c.s_off, S, b, c.Cin, c.Cout);
            P.demod_maxJ = std::max(P.demod_maxJ, c.Cout);
            P.demod_bwd_maxJ = std::max(P.demod_bwd_maxJ, c.Cin);
        }
        P.demod_tab = upload(ar, h0);
        P.demod_bwd_tab = upload(ar, h1);
    }
    if (ar.failed) return nullptr;
    SG2Plan* raw = pp.get();
    plans[b] = pp;
    return raw;
}

// mapping network: PixelNorm + n_mlp EqualLinear(fused_lrelu); h[0..n_mlp] are [b, sdim] buffers, h[n_mlp] = w
int SG2::run_mapping(int b, const float* z, float* zbuf, float* const* h, cudaStream_t st) {
    P2L_CUDA_CHECK(cudaMemcpyAsync(zbuf, z, (size_t)b * sdim * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (k_sg_mapping_fusable(b, sdim, n_mlp)) {   // one cluster launch for the whole network
        k_sg_mapping_fwd(zbuf, map_WT.data(), map_b.data(), map_scale, h, b, n_mlp, st);
        return 0;
    }
    k_pixelnorm_fwd(zbuf, h[0], b, sdim, st);
    for (int k = 0; k < n_mlp; ++k)
        k_fc_fwd(h[k], sdim, map_WT[k], map_b[k], map_scale, h[k + 1], sdim, b, sdim, sdim, 1, 0, st);
    return 0;
}

int SG2::style(int b, const float* z, float* w, cudaStream_t st) {
    if (!finalized) { set_error("sg2: style before finalize"); return -1; }
    // own scratch (not a full per-batch plan: this is called with thousands of samples for the latent statistics)
    auto it = style_scratch.find(b);
    if (it == style_scratch.end()) {
        float* p = style_ar.alloc<float>((size_t)(n_mlp + 2) * b * sdim);
        if (!p) return -1;
        it = style_scratch.emplace(b, p).first;
    }
    float* base = it->second;
    float* h[16];
    for (int k = 0; k <= n_mlp; ++k) h[k] = base + (size_t)(k + 1) * b * sdim;
    if (run_mapping(b, z, base, h, st)) return -1;
    P2L_CUDA_CHECK(cudaMemcpyAsync(w, h[n_mlp], (size_t)b * sdim * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int SG2::forward(int b, const float* z, const float* const* noise, float* img, cudaStream_t st) {
    if (!finalized) { set_error("sg2: forward before finalize"); return -1; }
    SG2Plan* Pp = plan(b);
    if (!Pp) return -1;
    SG2Plan& P = *Pp;
    if (run_mapping(b, z, P.z, P.h, st)) return -1;
    const float* w = P.h[n_mlp];
    // styles of every modulated conv / ToRGB (the same w feeds all of them in z search)
    k_fc_fwd(w, sdim, affT, aff_b, 1.f / std::sqrt((float)sdim), P.s_all, S, b, sdim, S, 0, 0, st);
    P.mode = 0;
    return synth(P, b, noise, img, st);
}

int SG2::forward_w(int b, const float* latent, const float* const* noise, float* img, cudaStream_t st) {
    if (!finalized) { set_error("sg2: forward_w before finalize"); return -1; }
    SG2Plan* Pp = plan(b);
    if (!Pp) return -1;
    SG2Plan& P = *Pp;
    const int nlat = n_latent(), ldl = nlat * sdim;
    const float sc = 1.f / std::sqrt((float)sdim);
    // every modulation reads ITS row of the latent: one small GEMV per layer on a column block of affT
    for (size_t l = 0; l < convs.size(); ++l) {
        const Conv& c = convs[l];
        k_fc_fwd_ld(latent + l * sdim, ldl, affT + c.s_off, S, aff_b + c.s_off, sc, P.s_all + c.s_off, S, b, sdim, c.Cin, 0, 0, st);
    }
    for (size_t t = 0; t < rgbs.size(); ++t) {
        const Rgb& r = rgbs[t];
        k_fc_fwd_ld(latent + (2 * t + 1) * sdim, ldl, affT + r.s_off, S, aff_b + r.s_off, sc, P.s_all + r.s_off, S, b, sdim, r.Cin, 0, 0, st);
    }
    P.mode = 1;
    return synth(P, b, noise, img, st);
}

int SG2::synth(SG2Plan& P, int b, const float* const* noise, float* img, cudaStream_t st) {
    const int nL = (int)convs.size();
    if (P.demod_tab) {
        k_sg_demod_batched(P.demod_tab, nL, b, P.demod_maxJ, st);
    } else {
        for (int l = 0; l < nL; ++l) {
            const Conv& c = convs[l];
            k_fc_fwd(P.s_all + c.s_off, S, c.wsqT, nullptr, 1.f, P.dm_all + c.dm_off, DM, b, c.Cin, c.Cout, 2, 1, st);
        }
    }
    // A_0 = const * s_0; every later A_l comes out of layer l-1's epilogue
    k_sg_modulate(const_in, 0, P.s_all + convs[0].s_off, S, P.A[0], b, convs[0].Hin, convs[0].Hin, convs[0].Cin, st);
    int t = 0;
    for (int l = 0; l < nL; ++l) {
        SG2Plan::Lay& q = P.L[l];
        q.f.p.sg_noise = noise ? noise[l] : nullptr;
        if (conv_op_launch(q.f, st)) return -1;
        if (has_rgb(l)) {
            const Rgb& r = rgbs[t];
            k_sg_weff(r.Wr, P.s_all + r.s_off, S, r.scale, P.weff[t], b, r.Cin, st);
            k_sg_torgb_fwd(q.x, P.weff[t], r.bias, t > 0 ? P.rgb[t - 1] : nullptr, P.rgb[t], b, r.H, r.H, r.Cin, st);
            ++t;
        }
    }
    // the noise images are needed again by the backward epilogues (they recompute the pre-activation from x)
    P.noise_ptrs.assign(nL, nullptr);
    if (noise)
        for (int l = 0; l < nL; ++l) P.noise_ptrs[l] = noise[l];
    const long n = (long)b * 3 * cfg.size * cfg.size;
    k_sg_clamp(P.rgb.back(), P.img, n, st);
    if (img && img != P.img) P2L_CUDA_CHECK(cudaMemcpyAsync(img, P.img, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    P.forward_done = true;
    return 0;
}

int SG2::synth_bwd(SG2Plan& P, int b, const float* dimg, float* const* dnoise, float out_scale, const float* row_scale,
                   cudaStream_t st) {
    const int T = (int)rgbs.size(), nL = (int)convs.size();
    P2L_CUDA_CHECK(cudaMemsetAsync(P.ds_all, 0, (size_t)b * S * sizeof(float), st));
    P2L_CUDA_CHECK(cudaMemsetAsync(P.ddm_all, 0, (size_t)b * DM * sizeof(float), st));
    for (int t = 0; t < T; ++t) P2L_CUDA_CHECK(cudaMemsetAsync(P.dweff[t], 0, (size_t)b * 3 * rgbs[t].Cin * sizeof(float), st));
    // gradient of every resolution's rgb image: d rgb_{t-1} = (adjoint of the FIR up-sampling)(d rgb_t)
    k_sg_clamp_bwd(P.rgb.back(), dimg, P.drgb[T - 1], (long)b * 3 * cfg.size * cfg.size, grad_scale(), st);
    for (int t = T - 1; t > 0; --t) k_sg_rgb_up_adjoint(P.drgb[t], P.drgb[t - 1], b, rgbs[t].H / 2, rgbs[t].H / 2, st);
    // gradient wrt x_l through its ToRGB branch -> dxrgb; style gradient of that ToRGB
    auto rgb_branch = [&](int l) {
        const int t = l / 2;   // l == 0 -> 0, l == 2t -> t
        const Rgb& r = rgbs[t];
        k_sg_torgb_bwd(P.drgb[t], P.L[l].x, P.weff[t], P.dxrgb, P.dweff[t], P.scratch, b, r.H, r.H, r.Cin, 0, st);
        k_sg_weff_bwd(P.dweff[t], r.Wr, r.scale, P.ds_all + r.s_off, S, b, r.Cin, st);
    };
    {   // last layer: its output feeds the last ToRGB only
        const int l = nL - 1;
        const Conv& c = convs[l];
        if (dnoise && dnoise[l]) {   // the noise gradient reads the gradient wrt x: keep it materialised
            rgb_branch(l);
            k_sg_noise_bwd(P.dxrgb, P.L[l].x, c.noise_w, dnoise[l], b, c.Hout, c.Hout, c.Cout, out_scale, row_scale, st);
            k_sg_post_bwd_x(P.dxrgb, P.L[l].x, P.dm_all + c.dm_off, DM, P.noise_ptrs[l], c.noise_w, c.bias, P.G[l & 1],
                            P.ddm_all + c.dm_off, P.scratch, b, c.Hout, c.Hout, c.Cout, st);
        } else {                     // one pass over x: ToRGB backward + the activation backward of the layer
            const int t = l / 2;
            const Rgb& r = rgbs[t];
            k_sg_torgb_post_bwd(P.drgb[t], P.L[l].x, P.weff[t], P.dweff[t], P.dm_all + c.dm_off, DM, P.noise_ptrs[l], c.noise_w, c.bias,
                                P.G[l & 1], P.ddm_all + c.dm_off, P.scratch, b, r.H, r.H, r.Cin, st);
            k_sg_weff_bwd(P.dweff[t], r.Wr, r.scale, P.ds_all + r.s_off, S, b, r.Cin, st);
        }
    }
    for (int l = nL - 1; l >= 1; --l) {
        const Conv& pv = convs[l - 1];
        SG2Plan::Lay& q = P.L[l];
        const bool rgb = has_rgb(l - 1);
        if (rgb) rgb_branch(l - 1);
        const bool want_dn = dnoise && dnoise[l - 1];
        q.d.p.sg_noise = P.noise_ptrs[l - 1];
        q.d.p.addin = rgb ? P.dxrgb : nullptr;
        q.d.p.dx2 = want_dn ? P.dx2 : nullptr;
        if (conv_op_launch(q.d, st)) return -1;
        if (want_dn) k_sg_noise_bwd(P.dx2, P.L[l - 1].x, pv.noise_w, dnoise[l - 1], b, pv.Hout, pv.Hout, pv.Cout, out_scale, row_scale, st);
    }
    {   // layer 0: ds_0 = sum dA_0 * const
        const Conv& c = convs[0];
        if (conv_op_launch(P.L[0].d, st)) return -1;
        k_sg_modulate_bwd(P.dA0, const_in, 0, P.s_all + c.s_off, S, nullptr, P.ds_all + c.s_off, S, P.scratch, b, c.Hin, c.Hin, c.Cin, st);
    }
    // (ds_l, ddm_{l-1}) from the dgrad epilogues' partial slots, then the gradient through every demodulation
    k_stat_reduce2(P.segs, P.nsegs, P.ds_all, S, P.ddm_all, DM, b, st);
    if (P.demod_bwd_tab) {
        k_sg_demod_bwd_batched(P.demod_bwd_tab, nL, b, P.demod_bwd_maxJ, st);
    } else {
        for (int l = 0; l < nL; ++l) {
            const Conv& c = convs[l];
            k_demod_bwd(P.ddm_all + c.dm_off, P.dm_all + c.dm_off, DM, P.s_all + c.s_off, S, c.wsq, P.ds_all + c.s_off, S, b, c.Cin, c.Cout, st);
        }
    }
    return 0;
}

int SG2::backward(int b, const float* dimg, float* dz, cudaStream_t st, float scale, const float* row_scale) {
    auto it = plans.find(b);
    if (it == plans.end() || !it->second->forward_done) { set_error("sg2: backward(b=%d) without a matching forward", b); return -1; }
    SG2Plan& P = *it->second;
    if (P.mode != 0) { set_error("sg2: backward (z search) after forward_w; use backward_w"); return -1; }
    if (synth_bwd(P, b, dimg, nullptr, 1.f, nullptr, st)) return -1;
    // styles -> w -> mapping network -> z
    k_fc_bwd(P.ds_all, S, nullptr, 0, aff, 1.f / std::sqrt((float)sdim), P.dw, sdim, b, sdim, S, 0, 0, st, b <= 24 ? P.scratch : nullptr);
    if (k_sg_mapping_fusable(b, sdim, n_mlp)) {
        k_sg_mapping_bwd(P.dw, map_W.data(), map_scale, P.h, P.z, dz, scale / grad_scale(), row_scale, b, n_mlp, st);
        return 0;
    }
    float *g = P.dw, *gn = P.g0;
    for (int k = n_mlp - 1; k >= 0; --k) {
        k_fc_bwd(g, sdim, P.h[k + 1], sdim, map_W[k], map_scale, gn, sdim, b, sdim, sdim, 1, 0, st);
        g = gn;
        gn = (gn == P.g0) ? P.g1 : P.g0;
    }
    k_pixelnorm_bwd(P.z, g, dz, b, sdim, scale / grad_scale(), row_scale, st);
    return 0;
}

int SG2::backward_w(int b, const float* dimg, float* dlatent, float* const* dnoise, cudaStream_t st, float scale,
                    const float* row_scale) {
    auto it = plans.find(b);
    if (it == plans.end() || !it->second->forward_done) { set_error("sg2: backward_w(b=%d) without a matching forward_w", b); return -1; }
    SG2Plan& P = *it->second;
    if (P.mode != 1) { set_error("sg2: backward_w after a z-search forward; use backward"); return -1; }
    const float out_scale = scale / grad_scale();
    if (synth_bwd(P, b, dimg, dnoise, out_scale, row_scale, st)) return -1;
    const int nlat = n_latent(), ldl = nlat * sdim;
    const float sc = 1.f / std::sqrt((float)sdim);
    P2L_CUDA_CHECK(cudaMemsetAsync(dlatent, 0, (size_t)b * ldl * sizeof(float), st));
    // styles -> their latent rows (row 2t+1 is shared by StyledConv 2t+1 and ToRGB t: accumulate)
    for (size_t l = 0; l < convs.size(); ++l) {
        const Conv& c = convs[l];
        k_fc_bwd(P.ds_all + c.s_off, S, nullptr, 0, aff + (size_t)c.s_off * sdim, sc, dlatent + l * sdim, ldl, b, sdim, c.Cin, 0, 1, st);
    }
    for (size_t t = 0; t < rgbs.size(); ++t) {
        const Rgb& r = rgbs[t];
        k_fc_bwd(P.ds_all + r.s_off, S, nullptr, 0, aff + (size_t)r.s_off * sdim, sc, dlatent + (2 * t + 1) * sdim, ldl, b, sdim, r.Cin, 0, 1, st);
    }
    k_sg_scale_out(dlatent, b, (long)ldl, out_scale, row_scale, st);
    return 0;
}

const float* SG2::last_image(int b) {
    auto it = plans.find(b);
    return it == plans.end() ? nullptr : it->second->img;
}

}  // namespace p2l
