// Native BigGAN-deep generator (see biggan.h). Layer algebra follows oracle/biggan.py, which
// restates pytorch_pretrained_biggan/model.py as reached via pix2latent/model/biggan.py:50-58.
//
// Fusion map (what the reference runs as separate ATen ops -> where it lives here):
//   conditional BN affine + ReLU (+ nearest x2)  -> epilogue of the PRODUCING convolution
//   residual add / channel drop / skip upsample  -> epilogue of conv_3
//   conv_to_rgb[:, :3] + tanh                    -> N=16 tile, 3 live channels, tanh epilogue
//   BN-affine gradients (sum dY, sum dY*Y)       -> epilogue reductions of the dgrad convolutions
//   weight gradients                              -> never computed (frozen generator)
#include "biggan.h"

#include <cmath>
#include <cstring>

namespace p2l {

static std::string lname(int idx, const char* rest) { return "generator.layers." + std::to_string(idx) + "." + rest; }

struct BigGANPlan {
    int b = 0;
    Arena ar;
    // latent side
    float *cond = nullptr, *a = nullptr, *s = nullptr, *S0 = nullptr, *S1 = nullptr, *G = nullptr, *dcond = nullptr,
          *dh0 = nullptr, *ones = nullptr;
    struct BB {
        act_t *in_raw, *in_act, *t1_lo, *t1, *t2, *t3, *out_raw, *out_act;
        ConvOp f[4], d[4];
        // BN-affine gradient partial sums of bn_0..bn_3 ([n][part][2][C], ConvGemmParams::statp)
        float* sp[4] = {nullptr, nullptr, nullptr, nullptr};
        int sp_parts[4] = {0, 0, 0, 0};
    };
    std::vector<BB> bb;
    StatSeg* segs = nullptr;  // device table for k_stat_reduce
    int nsegs = 0;
    // attention
    act_t *qkv = nullptr, *phi_p = nullptr, *phiT = nullptr, *g_p = nullptr, *gT = nullptr, *P = nullptr,
                  *O = nullptr, *attn_raw = nullptr, *attn_act = nullptr;
    unsigned char *idx_phi = nullptr, *idx_g = nullptr;
    float* S = nullptr;
    act_t *dO = nullptr, *dOT = nullptr, *dS = nullptr, *dST = nullptr, *PT = nullptr, *thetaT = nullptr,
                  *dqkv = nullptr, *dphi_p = nullptr, *dg_p = nullptr;
    ConvOp a_qkv, a_s, a_o, a_out, ad_out, ad_p, ad_theta, ad_phi, ad_g, ad_qkv;
    // "attn_fused": two-pass softmax in the S GEMM's epilogue, dS in the dP GEMM's epilogue (no fp32 logits in HBM);
    // "attn_emit_t" (needs attn_fused): theta^T, P^T, dO^T, dS^T come out of the producing epilogues (no transposes)
    bool attn_fused = false, attn_emit_t = false;
    ConvOp a_s1, a_s2, ad_pf;
    float *rowstat = nullptr, *Drow = nullptr;
    // image
    ConvOp f_rgb, d_rgb;
    act_t* col_rgb = nullptr;
    float* img = nullptr;  // internal copy target when the caller passes none
    float* rgbT = nullptr; // tap-expanded rgb head output [b, 27, R, R]
    // gradient ping-pong
    act_t *dhA = nullptr, *dhB = nullptr, *g1 = nullptr, *g2 = nullptr, *g3 = nullptr, *dh_pool = nullptr;
    double flops_fwd = 0, flops_bwd = 0;
    int launches_fwd = 0, launches_bwd = 0;
    bool forward_done = false;
};

BigGAN::~BigGAN() {}

// ----------------------------------------------------------------------------- finalize
int BigGAN::finalize() {
    if (finalized) return 0;
    const int ch = cfg.channel_width;
    cdim = cfg.z_dim + cfg.class_embed_dim;
    const int nL = cfg.n_layers;
    if (nL <= 0 || nL > P2L_MAX_LAYERS) { set_error("bad n_layers %d", nL); return -1; }
    // BN statistics row (pytorch_pretrained_biggan BigGANBatchNorm.forward)
    const double step = 1.0 / (cfg.n_stats - 1);
    double ip;
    const double coef = std::modf((double)cfg.truncation / step, &ip);
    const int start = (int)ip;

    // ---- enumerate blocks and BNs
    int H = 4;
    blocks.clear();
    bns.clear();
    int off = 0;
    for (int i = 0; i < nL; ++i) {
        Block bl{};
        bl.in = ch * cfg.in_mult[i];
        bl.out = ch * cfg.out_mult[i];
        bl.mid = bl.in / 4;
        bl.up = cfg.up[i] != 0;
        bl.Hin = H;
        bl.Hout = bl.up ? 2 * H : H;
        if (bl.mid % 64 || bl.out % 64) { set_error("block %d: channel counts must be multiples of 64", i); return -1; }
        if (bl.in != bl.out && bl.out * 2 != bl.in) { set_error("block %d: out must equal in or in/2", i); return -1; }
        const int cs[4] = {bl.in, bl.mid, bl.mid, bl.mid};
        for (int k = 0; k < 4; ++k) {
            bl.bn[k] = (int)bns.size();
            bns.push_back({cs[k], off, true});
            off += cs[k];
        }
        H = bl.Hout;
        blocks.push_back(bl);
    }
    C_cond = off;
    C_last = blocks.back().out;
    H_out = H;
    final_bn = (int)bns.size();
    bns.push_back({C_last, off, false});
    C_all = off + C_last;
    C0 = blocks[0].in;
    genz_J = 16 * C0;

    // ---- BN tables + cond linears
    std::vector<float> h_mean(C_all), h_istd(C_all), h_ws((size_t)C_cond * cdim), h_wo((size_t)C_cond * cdim);
    auto fill_stats = [&](const std::string& prefix, int C, int o) -> int {
        const auto* rm = stage.get(prefix + ".running_means", (long)cfg.n_stats * C);
        const auto* rv = stage.get(prefix + ".running_vars", (long)cfg.n_stats * C);
        if (!rm || !rv) return -1;
        for (int c = 0; c < C; ++c) {
            double m, v;
            if (coef != 0.0) {
                m = (*rm)[(size_t)start * C + c] * coef + (*rm)[(size_t)(start + 1) * C + c] * (1 - coef);
                v = (*rv)[(size_t)start * C + c] * coef + (*rv)[(size_t)(start + 1) * C + c] * (1 - coef);
            } else {
                m = (*rm)[(size_t)start * C + c];
                v = (*rv)[(size_t)start * C + c];
            }
            h_mean[o + c] = (float)m;
            h_istd[o + c] = (float)(1.0 / std::sqrt(v + (double)cfg.eps));
        }
        return 0;
    };
    for (int i = 0; i < nL; ++i) {
        const int li = i + ((cfg.attention_pos >= 0 && i >= cfg.attention_pos) ? 1 : 0);
        for (int k = 0; k < 4; ++k) {
            const BN& bn = bns[blocks[i].bn[k]];
            const std::string pre = lname(li, ("bn_" + std::to_string(k)).c_str());
            if (fill_stats(pre, bn.C, bn.off)) return -1;
            const auto* sw = stage.get(pre + ".scale.weight", (long)bn.C * cdim);
            const auto* ow = stage.get(pre + ".offset.weight", (long)bn.C * cdim);
            if (!sw || !ow) return -1;
            std::memcpy(&h_ws[(size_t)bn.off * cdim], sw->data(), sw->size() * sizeof(float));
            std::memcpy(&h_wo[(size_t)bn.off * cdim], ow->data(), ow->size() * sizeof(float));
        }
    }
    if (fill_stats("generator.bn", C_last, C_cond)) return -1;
    mean = upload(weights, h_mean);
    inv_std = upload(weights, h_istd);
    {   // cond -> (gain, offset) as ONE GEMV + bias with the BN statistics folded in:
        //   a = (1 + cond.Ws) * istd            = istd          + cond . (Ws * istd)
        //   s = cond.Wo - mean * a              = -mean * istd  + cond . (Wo - mean * istd * Ws)
        // transposed [cdim][2*C_cond] so a warp of channels reads contiguous rows
        if (cdim % 128) { set_error("biggan: z_dim + class_embed_dim must be a multiple of 128"); return -1; }
        std::vector<float> t((size_t)2 * C_cond * cdim), bias((size_t)2 * C_cond);
        for (int c = 0; c < C_cond; ++c) {
            const double is = h_istd[c], m = h_mean[c];
            bias[c] = (float)is;
            bias[(size_t)C_cond + c] = (float)(-m * is);
            for (int k = 0; k < cdim; ++k) {
                const double ws = h_ws[(size_t)c * cdim + k], wo = h_wo[(size_t)c * cdim + k];
                t[(size_t)k * 2 * C_cond + c] = (float)(ws * is);
                t[(size_t)k * 2 * C_cond + C_cond + c] = (float)(wo - m * is * ws);
            }
        }
        WT_as = upload(weights, t);
        bias_as = upload(weights, bias);
    }
    {
        std::vector<float> cat((size_t)2 * C_cond * cdim);
        std::memcpy(cat.data(), h_ws.data(), h_ws.size() * sizeof(float));
        std::memcpy(cat.data() + h_ws.size(), h_wo.data(), h_wo.size() * sizeof(float));
        Wcat = upload(weights, cat);
    }
    {
        const auto* w = stage.get("generator.bn.weight", C_last);
        const auto* bsv = stage.get("generator.bn.bias", C_last);
        if (!w || !bsv) return -1;
        unc_weight = upload(weights, *w);
        unc_bias = upload(weights, *bsv);
    }
    {
        const auto* w = stage.get("generator.gen_z.weight", (long)genz_J * cdim);
        const auto* bsv = stage.get("generator.gen_z.bias", genz_J);
        if (!w || !bsv) return -1;
        genz_W = upload(weights, *w);  // row-major [J][cdim] for the backward (dcond) pass
        std::vector<float> t(w->size());
        for (int j = 0; j < genz_J; ++j)
            for (int k = 0; k < cdim; ++k) t[(size_t)k * genz_J + j] = (*w)[(size_t)j * cdim + k];
        genz_WT = upload(weights, t);
        genz_b = upload(weights, *bsv);
    }
    // ---- conv weights
    for (int i = 0; i < nL; ++i) {
        Block& bl = blocks[i];
        const int li = i + ((cfg.attention_pos >= 0 && i >= cfg.attention_pos) ? 1 : 0);
        const int ci[4] = {bl.in, bl.mid, bl.mid, bl.mid};
        const int co[4] = {bl.mid, bl.mid, bl.mid, bl.out};
        const int ks[4] = {1, 3, 3, 1};
        for (int k = 0; k < 4; ++k) {
            const std::string pre = lname(li, ("conv_" + std::to_string(k)).c_str());
            const auto* w = stage.get(pre + ".weight", (long)co[k] * ci[k] * ks[k] * ks[k]);
            const auto* bsv = stage.get(pre + ".bias", co[k]);
            if (!w || !bsv) return -1;
            bl.w[k] = upload(weights, pack_conv_fwd(*w, co[k], ci[k], ks[k], ks[k]));
            bl.wt[k] = upload(weights, pack_conv_dgrad(*w, co[k], ci[k], ks[k], ks[k]));
            bl.bias[k] = upload(weights, *bsv);
        }
    }
    // ---- attention
    if (cfg.attention_pos >= 0) {
        if (cfg.attention_pos >= nL) { set_error("attention_pos out of range"); return -1; }
        const int li = cfg.attention_pos;
        attn.C = blocks[cfg.attention_pos].in;
        attn.H = blocks[cfg.attention_pos].Hin;
        attn.dq = attn.C / 8;
        attn.dv = attn.C / 2;
        if (cfg.attention_pos < 1 || attn.dq % 64 || (attn.H / 2) * (attn.H / 2) < 128 || (attn.H * attn.H / 4) % 64) {
            set_error("attention shape unsupported (C=%d, H=%d)", attn.C, attn.H);
            return -1;
        }
        const auto* wt = stage.get(lname(li, "snconv1x1_theta.weight"), (long)attn.dq * attn.C);
        const auto* wp = stage.get(lname(li, "snconv1x1_phi.weight"), (long)attn.dq * attn.C);
        const auto* wg = stage.get(lname(li, "snconv1x1_g.weight"), (long)attn.dv * attn.C);
        const auto* wo = stage.get(lname(li, "snconv1x1_o_conv.weight"), (long)attn.C * attn.dv);
        const auto* gm = stage.get(lname(li, "gamma"), 1);
        if (!wt || !wp || !wg || !wo || !gm) return -1;
        const int nq = 2 * attn.dq + attn.dv;
        std::vector<float> cat((size_t)nq * attn.C);
        std::memcpy(cat.data(), wt->data(), wt->size() * 4);
        std::memcpy(cat.data() + wt->size(), wp->data(), wp->size() * 4);
        std::memcpy(cat.data() + wt->size() + wp->size(), wg->data(), wg->size() * 4);
        attn.wqkv = upload(weights, pack_conv_fwd(cat, nq, attn.C, 1, 1));
        attn.wqkv_t = upload(weights, pack_conv_dgrad(cat, nq, attn.C, 1, 1));
        attn.wo = upload(weights, pack_conv_fwd(*wo, attn.C, attn.dv, 1, 1));
        attn.wo_t = upload(weights, pack_conv_dgrad(*wo, attn.C, attn.dv, 1, 1));
        attn.gamma = upload(weights, *gm);
    }
    // ---- rgb: only the 3 live output channels of conv_to_rgb (biggan slices [:, :3])
    {
        const auto* w = stage.get("generator.conv_to_rgb.weight", (long)C_last * C_last * 9);
        const auto* bsv = stage.get("generator.conv_to_rgb.bias", C_last);
        if (!w || !bsv) return -1;
        // The 3-channel 3x3 head as a 1x1 GEMM with N = 27 = (tap, channel) outputs + a 9-tap gather:
        // a tcgen05.mma costs the same ~100 cycles for N = 16 as for N = 64, so nine K=128 taps with
        // N = 3 (72 MMAs per tile) are replaced by one K = 128 pass with N = 27 (8 MMAs per tile).
        {
            std::vector<act_t> t((size_t)27 * C_last);
            for (int tap = 0; tap < 9; ++tap)
                for (int o = 0; o < 3; ++o)
                    for (int c = 0; c < C_last; ++c)
                        t[((size_t)tap * 3 + o) * C_last + c] = host_f2bf((*w)[(((size_t)o * C_last + c) * 3 + tap / 3) * 3 + tap % 3]);
            wrgb = upload(weights, t);
        }
        std::vector<float> b3(bsv->begin(), bsv->begin() + 3);
        brgb = upload(weights, b3);
        // dgrad operand for the im2col'd gradient: [C_last][64], k = (r*3+s)*3 + o -> W[o, c, r, s]
        std::vector<act_t> t((size_t)C_last * 64, host_f2bf(0.f));
        for (int c = 0; c < C_last; ++c)
            for (int r = 0; r < 3; ++r)
                for (int s2 = 0; s2 < 3; ++s2)
                    for (int o = 0; o < 3; ++o)
                        t[(size_t)c * 64 + (r * 3 + s2) * 3 + o] = host_f2bf((*w)[(((size_t)o * C_last + c) * 3 + r) * 3 + s2]);
        wrgb_t = upload(weights, t);
    }
    if (weights.failed) return -1;
    stage.t.clear();
    finalized = true;
    return 0;
}

// ----------------------------------------------------------------------------- plan
static int pick_bn_for(int Cout, long m_tiles, long K) {
    if (Cout <= 16) return 16;
    const int cands[3] = {256, 128, 64};
    // small-K launches are epilogue-bound: N <= 128 tiles run two CTAs per SM and (K <= 512) store
    // through TMA; N = 256 tiles only pay off when the main loop is long
    for (int k = (K <= 1024 ? 1 : 0); k < 3; ++k)
        if (Cout % cands[k] == 0 && m_tiles * (Cout / cands[k]) >= num_sms()) return cands[k];
    for (int k = 2; k >= 0; --k)
        if (Cout % cands[k] == 0) return cands[k];
    return 64;
}
static long m_tiles_for(int b, int H, int W) {
    int tw = 1; while (tw < W) tw <<= 1; if (tw > 16) tw = 16;
    int th = 1; while (th < H) th <<= 1; if (th > 128 / tw) th = 128 / tw;
    const int nb = 128 / (tw * th);
    return (long)((W + tw - 1) / tw) * ((H + th - 1) / th) * ((b + nb - 1) / nb);
}

struct OpB {  // small builder
    ConvDesc d;
    OpB(const void* A, int N, int H, int W, int C, int c0, int Cin, const void* B, int Cout, int k, int mode) {
        d.A = A; d.A_N = N; d.A_H = H; d.A_W = W; d.A_C = C; d.a_c0 = c0; d.Cin = Cin;
        d.B = B; d.Cout = Cout; d.kh = d.kw = k; d.pad_h = d.pad_w = k / 2;
        d.NI = N; d.H = H; d.W = W; d.mode = mode;
        d.BN = pick_bn_for(Cout, m_tiles_for(N, H, W), (long)k * k * Cin);
    }
};

BigGANPlan* BigGAN::plan(int b) {
    auto it = plans.find(b);
    if (it != plans.end()) return it->second.get();
    std::shared_ptr<BigGANPlan> pp(new BigGANPlan());
    BigGANPlan& P = *pp;
    P.b = b;
    Arena& ar = P.ar;
    typedef act_t bf;
    P.cond = ar.alloc<float>((size_t)b * cdim);
    P.a = ar.alloc<float>((size_t)b * C_all);
    P.s = ar.alloc<float>((size_t)b * C_all);
    P.S0 = ar.alloc<float>((size_t)b * C_all, true);
    P.S1 = ar.alloc<float>((size_t)b * C_all, true);
    P.G = ar.alloc<float>((size_t)b * 2 * C_cond);
    P.dcond = ar.alloc<float>((size_t)(k_dcond_blocks(2 * C_cond) + k_dcond_blocks(genz_J)) * b * cdim);  // partial sums per row block
    P.dh0 = ar.alloc<float>((size_t)b * genz_J);
    const int nL = (int)blocks.size();
    P.bb.resize(nL);
    size_t max_dh = 0, max_g = 0;
    // activations
    act_t *cur_raw = ar.alloc<bf>((size_t)b * genz_J), *cur_act = ar.alloc<bf>((size_t)b * genz_J);
    for (int i = 0; i < nL; ++i) {
        const Block& bl = blocks[i];
        BigGANPlan::BB& B = P.bb[i];
        const size_t pin = (size_t)b * bl.Hin * bl.Hin, pout = (size_t)b * bl.Hout * bl.Hout;
        if (attn.C && i == cfg.attention_pos) {
            // attention sits between the previous block's raw output and this block's input
            P.attn_raw = ar.alloc<bf>(pin * bl.in);
            P.attn_act = ar.alloc<bf>(pin * bl.in);
            B.in_raw = P.attn_raw;
            B.in_act = P.attn_act;
        } else {
            B.in_raw = cur_raw;
            B.in_act = cur_act;
        }
        B.t1_lo = ar.alloc<bf>(pin * bl.mid);
        B.t1 = bl.up ? ar.alloc<bf>(pout * bl.mid) : B.t1_lo;
        B.t2 = ar.alloc<bf>(pout * bl.mid);
        B.t3 = ar.alloc<bf>(pout * bl.mid);
        B.out_raw = ar.alloc<bf>(pout * bl.out);
        B.out_act = ar.alloc<bf>(pout * bl.out);
        cur_raw = B.out_raw;
        cur_act = B.out_act;
        max_dh = std::max(max_dh, std::max(pout * bl.out, pin * bl.in));
        max_g = std::max(max_g, pout * bl.mid);
        // BN-gradient partial buffers: bn_0 is filled by d0 (input grid), bn_1 by d1 (or the pooling kernel of an up
        // block, input grid), bn_2 / bn_3 by d2 / d3 (output grid)
        const int cs[4] = {bl.in, bl.mid, bl.mid, bl.mid};
        B.sp_parts[0] = conv_stat_parts_max(bl.Hin, bl.Hin);
        B.sp_parts[1] = bl.up ? k_pool_bnrelu_parts(bl.Hin, bl.Hin) : conv_stat_parts_max(bl.Hout, bl.Hout);
        B.sp_parts[2] = B.sp_parts[3] = conv_stat_parts_max(bl.Hout, bl.Hout);
        for (int k = 0; k < 4; ++k) B.sp[k] = ar.alloc<float>((size_t)b * B.sp_parts[k] * 2 * cs[k]);
    }
    P.dhA = ar.alloc<bf>(max_dh);
    P.dhB = ar.alloc<bf>(max_dh);
    P.g1 = ar.alloc<bf>(max_g);
    P.g2 = ar.alloc<bf>(max_g);
    P.g3 = ar.alloc<bf>(max_g);
    P.dh_pool = ar.alloc<bf>(max_dh / 4 + 64);  // 2x2-pooled skip gradient of up blocks
    const int R = H_out;
    P.col_rgb = ar.alloc<bf>((size_t)b * R * R * 64);
    P.img = ar.alloc<float>((size_t)b * 3 * R * R);
    P.rgbT = ar.alloc<float>((size_t)b * 27 * R * R);
    if (ar.failed) return nullptr;

    auto build = [&](ConvOp* op, OpB& ob, double* flops, int* launches) -> int {
        if (conv_op_build(op, ob.d)) return -1;
        *flops += op->flops;
        *launches += 1;
        return 0;
    };
    std::vector<StatSeg> segs;
    // ---- per-block ops
    for (int i = 0; i < nL; ++i) {
        const Block& bl = blocks[i];
        BigGANPlan::BB& B = P.bb[i];
        const int Hi = bl.Hin, Ho = bl.Hout;
        const bool attn_next = attn.C && (i + 1 == cfg.attention_pos);
        const int next_bn = (i + 1 < nL) ? blocks[i + 1].bn[0] : final_bn;
        act_t* dh_out = ((nL - 1 - i) % 2 == 0) ? P.dhA : P.dhB;  // dh_out lives in dhA when (nL-1-i) is even, dh_in goes to the other
        act_t* dh_in = ((nL - 1 - i) % 2 == 0) ? P.dhB : P.dhA;
        int parts_used[4] = {0, 0, 0, 0};
        const float *aff_a = P.a, *aff_s = P.s;
        auto stat = [&](ConvGemmParams& e, int k, int C) {
            e.statp = B.sp[k]; e.statp_parts = B.sp_parts[k]; e.statp_C = C;
        };
        {   // f0: 1x1 in->mid on in_act; epilogue bn_1+relu (+x2 replicate)
            OpB o(B.in_act, b, Hi, Hi, bl.in, 0, bl.in, bl.w[0], bl.mid, 1, EPI_FWD);
            ConvGemmParams& e = o.d.epi;
            e.bias = bl.bias[0];
            e.aff_a = aff_a + bns[bl.bn[1]].off; e.aff_s = aff_s + bns[bl.bn[1]].off; e.aff_stride = C_all; e.relu = 1;
            e.act = B.t1; e.act_C = bl.mid; e.act_up = bl.up ? 1 : 0;
            e.act_lo = bl.up ? B.t1_lo : nullptr;
            if (build(&B.f[0], o, &P.flops_fwd, &P.launches_fwd)) return nullptr;
        }
        {   // f1: 3x3 mid->mid
            OpB o(B.t1, b, Ho, Ho, bl.mid, 0, bl.mid, bl.w[1], bl.mid, 3, EPI_FWD);
            ConvGemmParams& e = o.d.epi;
            e.bias = bl.bias[1];
            e.aff_a = aff_a + bns[bl.bn[2]].off; e.aff_s = aff_s + bns[bl.bn[2]].off; e.aff_stride = C_all; e.relu = 1;
            e.act = B.t2; e.act_C = bl.mid;
            if (build(&B.f[1], o, &P.flops_fwd, &P.launches_fwd)) return nullptr;
        }
        {   // f2
            OpB o(B.t2, b, Ho, Ho, bl.mid, 0, bl.mid, bl.w[2], bl.mid, 3, EPI_FWD);
            ConvGemmParams& e = o.d.epi;
            e.bias = bl.bias[2];
            e.aff_a = aff_a + bns[bl.bn[3]].off; e.aff_s = aff_s + bns[bl.bn[3]].off; e.aff_stride = C_all; e.relu = 1;
            e.act = B.t3; e.act_C = bl.mid;
            if (build(&B.f[2], o, &P.flops_fwd, &P.launches_fwd)) return nullptr;
        }
        {   // f3: 1x1 mid->out + skip; raw; next BN + relu
            OpB o(B.t3, b, Ho, Ho, bl.mid, 0, bl.mid, bl.w[3], bl.out, 1, EPI_FWD);
            ConvGemmParams& e = o.d.epi;
            e.bias = bl.bias[3];
            e.resid = B.in_raw; e.resid_C = bl.in; e.resid_shift = bl.up ? 1 : 0;
            // the raw (pre-BN) output feeds the next block's skip / the attention; after the last block
            // nothing reads it
            if (i + 1 < nL) { e.raw = B.out_raw; e.raw_C = bl.out; }
            if (!attn_next) {
                e.aff_a = aff_a + bns[next_bn].off; e.aff_s = aff_s + bns[next_bn].off; e.aff_stride = C_all; e.relu = 1;
                e.act = B.out_act; e.act_C = bl.out;
            }
            if (build(&B.f[3], o, &P.flops_fwd, &P.launches_fwd)) return nullptr;
        }
        // ---- backward ops
        {   // d3: dh_out -> g3 (through bn_3/relu)
            OpB o(dh_out, b, Ho, Ho, bl.out, 0, bl.out, bl.wt[3], bl.mid, 1, EPI_BWD);
            ConvGemmParams& e = o.d.epi;
            e.saved = B.t3; e.saved_C = bl.mid;
            stat(e, 3, bl.mid);
            e.aff_a = aff_a + bns[bl.bn[3]].off; e.aff_stride = C_all;
            e.dx = P.g3; e.dx_C = bl.mid;
            if (build(&B.d[3], o, &P.flops_bwd, &P.launches_bwd)) return nullptr;
            parts_used[3] = B.d[3].stat_parts;
        }
        {   // d2: g3 -> g2 (through bn_2/relu)
            OpB o(P.g3, b, Ho, Ho, bl.mid, 0, bl.mid, bl.wt[2], bl.mid, 3, EPI_BWD);
            ConvGemmParams& e = o.d.epi;
            e.saved = B.t2; e.saved_C = bl.mid;
            stat(e, 2, bl.mid);
            e.aff_a = aff_a + bns[bl.bn[2]].off; e.aff_stride = C_all;
            e.dx = P.g2; e.dx_C = bl.mid;
            if (build(&B.d[2], o, &P.flops_bwd, &P.launches_bwd)) return nullptr;
            parts_used[2] = B.d[2].stat_parts;
        }
        {   // d1: g2 -> g1 (bn_1/relu; through the x2 upsample when bl.up: plain dgrad into g3, pooled later)
            OpB o(P.g2, b, Ho, Ho, bl.mid, 0, bl.mid, bl.wt[1], bl.mid, 3, EPI_BWD);
            ConvGemmParams& e = o.d.epi;
            if (!bl.up) {
                e.saved = B.t1_lo; e.saved_C = bl.mid;
                stat(e, 1, bl.mid);
                e.aff_a = aff_a + bns[bl.bn[1]].off; e.aff_stride = C_all;
                e.dx = P.g1; e.dx_C = bl.mid;
            } else {
                e.dx = P.g3; e.dx_C = bl.mid;  // g3 is free again: holds the hi-res gradient
                P.launches_bwd += 1;           // + k_pool_bnrelu_bwd
            }
            if (build(&B.d[1], o, &P.flops_bwd, &P.launches_bwd)) return nullptr;
            parts_used[1] = bl.up ? B.sp_parts[1] : B.d[1].stat_parts;
        }
        {   // d0: g1 -> dh_in (bn_0/relu) + skip gradient
            OpB o(P.g1, b, Hi, Hi, bl.mid, 0, bl.mid, bl.wt[0], bl.in, 1, EPI_BWD);
            ConvGemmParams& e = o.d.epi;
            e.saved = B.in_act; e.saved_C = bl.in;
            stat(e, 0, bl.in);
            e.aff_a = aff_a + bns[bl.bn[0]].off; e.aff_stride = C_all;
            // skip gradient: dh_out itself, or (up block) its 2x2-pooled copy written by k_pool2x2_sum
            e.addin = bl.up ? P.dh_pool : dh_out;
            e.addin_C = bl.out; e.addin_climit = bl.out; e.addin_pool = 0;
            e.dx = dh_in; e.dx_C = bl.in;
            if (i == 0) { e.dx_f32 = P.dh0; e.dx_f32_C = bl.in; }
            if (build(&B.d[0], o, &P.flops_bwd, &P.launches_bwd)) return nullptr;
            parts_used[0] = B.d[0].stat_parts;
        }
        const int cs[4] = {bl.in, bl.mid, bl.mid, bl.mid};
        for (int k = 0; k < 4; ++k)
            for (int c0 = 0; c0 < cs[k]; c0 += 32) segs.push_back({B.sp[k], parts_used[k], B.sp_parts[k], cs[k], bns[bl.bn[k]].off, c0, bns[bl.bn[k]].off});
    }
    P.nsegs = (int)segs.size();
    P.segs = upload(ar, segs);
    if (ar.failed) return nullptr;
    // ---- attention ops
    if (attn.C) {
        const int ap = cfg.attention_pos;
        const int H = attn.H, C = attn.C, dq = attn.dq, dv = attn.dv, nq = 2 * dq + dv;
        const int Nq = H * H, Nk = Nq / 4;
        const size_t px = (size_t)b * Nq;
        P.qkv = ar.alloc<bf>(px * nq);
        P.phi_p = ar.alloc<bf>((size_t)b * Nk * dq);
        P.phiT = ar.alloc<bf>((size_t)b * Nk * dq);
        P.g_p = ar.alloc<bf>((size_t)b * Nk * dv);
        P.gT = ar.alloc<bf>((size_t)b * Nk * dv);
        P.idx_phi = ar.alloc<unsigned char>((size_t)b * Nk * dq);
        P.idx_g = ar.alloc<unsigned char>((size_t)b * Nk * dv);
        P.S = ar.alloc<float>(px * Nk);
        P.P = ar.alloc<bf>(px * Nk);
        P.O = ar.alloc<bf>(px * dv);
        P.dO = ar.alloc<bf>(px * dv);
        P.dOT = ar.alloc<bf>(px * dv);
        P.dS = ar.alloc<bf>(px * Nk);
        P.dST = ar.alloc<bf>(px * Nk);
        P.PT = ar.alloc<bf>(px * Nk);
        P.thetaT = ar.alloc<bf>(px * dq);
        P.dqkv = ar.alloc<bf>(px * nq);
        P.dphi_p = ar.alloc<bf>((size_t)b * Nk * dq);
        P.dg_p = ar.alloc<bf>((size_t)b * Nk * dv);
        if (ar.failed) return nullptr;
        const act_t* x_raw = P.bb[ap - 1].out_raw;  // attention input (previous block's raw output)
        const int Hk = H / 2;
        P.launches_fwd += 3;  // 2 pools + softmax
        P.launches_bwd += 8;  // transposes x4, softmax bwd, pool bwd x2 ... (counted below as launched)
        P.attn_fused = get_option("attn_fused") != 0;
        // the transposed operands come out of the TMA-I/O (qkv, dO) and row-fusion (P, dS) kernels: K <= tma_kmax there
        P.attn_emit_t = P.attn_fused && get_option("attn_emit_t") != 0 && get_option("tma_out") != 0 &&
                        C <= get_option("tma_kmax") && nq % 64 == 0 && dv % 64 == 0;
        {   // qkv = x W_qkv^T
            OpB o(x_raw, b, H, H, C, 0, C, attn.wqkv, nq, 1, EPI_FWD);
            o.d.epi.raw = P.qkv; o.d.epi.raw_C = nq;
            if (P.attn_emit_t) { o.d.epi.outT = P.thetaT; o.d.epi.outT_c0 = 0; o.d.epi.outT_c1 = dq; }
            if (build(&P.a_qkv, o, &P.flops_fwd, &P.launches_fwd)) return nullptr;
        }
        {   // S = theta phi_p^T (fp32)
            OpB o(P.qkv, b, H, H, nq, 0, dq, P.phi_p, Nk, 1, EPI_FWD);
            o.d.B_batch = b; o.d.epi.raw_f32 = P.S; o.d.epi.raw_f32_C = Nk;
            if (build(&P.a_s, o, &P.flops_fwd, &P.launches_fwd)) return nullptr;
        }
        if (P.attn_fused) {
            // pass 1: per-row (max, sum exp) of every N tile of S = theta phi_p^T; pass 2: P = exp(S - M) / L, 16-bit.
            // The fp32 logits never reach HBM (2 x 302 MB per step at the bench shape) and k_softmax_fwd goes away.
            OpB o1(P.qkv, b, H, H, nq, 0, dq, P.phi_p, Nk, 1, EPI_FWD);
            o1.d.B_batch = b;
            // the row statistics are combined from one (max, sum) pair per N tile: the tile width must not depend on the
            // batch size, or a candidate's softmax would round differently in a smaller launch (sharded == unsharded)
            o1.d.BN = (Nk % 128 == 0) ? 128 : 64;
            const int nt = (Nk + o1.d.BN - 1) / o1.d.BN;
            P.rowstat = ar.alloc<float>(px * nt * 2);
            P.Drow = ar.alloc<float>(px);
            if (ar.failed) return nullptr;
            o1.d.epi.rowstat = P.rowstat; o1.d.epi.rowstat_nt = nt;
            if (build(&P.a_s1, o1, &P.flops_fwd, &P.launches_fwd)) return nullptr;
            OpB o2(P.qkv, b, H, H, nq, 0, dq, P.phi_p, Nk, 1, EPI_FWD);
            o2.d.B_batch = b; o2.d.BN = o1.d.BN;
            o2.d.epi.rowstat_in = P.rowstat; o2.d.epi.rowstat_nt = nt;
            o2.d.epi.raw = P.P; o2.d.epi.raw_C = Nk;
            if (P.attn_emit_t) { o2.d.epi.outT = P.PT; o2.d.epi.outT_c0 = 0; o2.d.epi.outT_c1 = Nk; }
            if (build(&P.a_s2, o2, &P.flops_fwd, &P.launches_fwd)) return nullptr;
            // dS = P o (dO g_p^T - rowsum(dO o O)) straight out of the dP GEMM (k_softmax_bwd and the fp32 dP go away)
            OpB o3(P.dO, b, H, H, dv, 0, dv, P.g_p, Nk, 1, EPI_FWD);
            o3.d.B_batch = b;
            o3.d.epi.rowsub = P.Drow; o3.d.epi.mulin = P.P; o3.d.epi.mulin_C = Nk;
            o3.d.epi.raw = P.dS; o3.d.epi.raw_C = Nk;
            if (P.attn_emit_t) { o3.d.epi.outT = P.dST; o3.d.epi.outT_c0 = 0; o3.d.epi.outT_c1 = Nk; }
            if (build(&P.ad_pf, o3, &P.flops_bwd, &P.launches_bwd)) return nullptr;
        }
        {   // O = P g_p
            OpB o(P.P, b, H, H, Nk, 0, Nk, P.gT, dv, 1, EPI_FWD);
            o.d.B_batch = b; o.d.epi.raw = P.O; o.d.epi.raw_C = dv;
            if (build(&P.a_o, o, &P.flops_fwd, &P.launches_fwd)) return nullptr;
        }
        {   // out = x + gamma * O W_o^T ; act = relu(bn_0(next block))
            OpB o(P.O, b, H, H, dv, 0, dv, attn.wo, C, 1, EPI_FWD);
            ConvGemmParams& e = o.d.epi;
            e.alpha_ptr = attn.gamma;
            e.resid = x_raw; e.resid_C = C; e.resid_shift = 0;
            e.raw = P.attn_raw; e.raw_C = C;
            const int nb0 = blocks[ap].bn[0];
            e.aff_a = P.a + bns[nb0].off; e.aff_s = P.s + bns[nb0].off; e.aff_stride = C_all; e.relu = 1;
            e.act = P.attn_act; e.act_C = C;
            if (build(&P.a_out, o, &P.flops_fwd, &P.launches_fwd)) return nullptr;
        }
        // backward: gradient wrt attention output arrives in the buffer block `ap` wrote as dh_in
        act_t* dh_attn_out = ((nL - 1 - ap) % 2 == 0) ? P.dhB : P.dhA;
        // block ap-1 reads its dh_out from the same ping-pong buffer block ap wrote dh_in to, so the
        // attention input gradient is produced in place (each thread reads its addin elements
        // before overwriting exactly those elements).
        act_t* dh_attn_in = dh_attn_out;
        {   // dO = gamma * dh W_o
            OpB o(dh_attn_out, b, H, H, C, 0, C, attn.wo_t, dv, 1, EPI_BWD);
            o.d.epi.alpha_ptr = attn.gamma; o.d.epi.dx = P.dO; o.d.epi.dx_C = dv;
            if (P.attn_emit_t) { o.d.epi.outT = P.dOT; o.d.epi.outT_c0 = 0; o.d.epi.outT_c1 = dv; }
            if (build(&P.ad_out, o, &P.flops_bwd, &P.launches_bwd)) return nullptr;
        }
        {   // dP = dO g_p^T (fp32, reuses S)
            OpB o(P.dO, b, H, H, dv, 0, dv, P.g_p, Nk, 1, EPI_FWD);
            o.d.B_batch = b; o.d.epi.raw_f32 = P.S; o.d.epi.raw_f32_C = Nk;
            if (build(&P.ad_p, o, &P.flops_bwd, &P.launches_bwd)) return nullptr;
        }
        {   // dtheta = dS phi_p  -> dqkv[:, 0:dq]
            OpB o(P.dS, b, H, H, Nk, 0, Nk, P.phiT, dq, 1, EPI_FWD);
            o.d.B_batch = b; o.d.epi.raw = P.dqkv; o.d.epi.raw_C = nq;
            if (build(&P.ad_theta, o, &P.flops_bwd, &P.launches_bwd)) return nullptr;
        }
        {   // dphi_p = dS^T theta
            OpB o(P.dST, b, Hk, Hk, Nq, 0, Nq, P.thetaT, dq, 1, EPI_FWD);
            o.d.B_batch = b; o.d.epi.raw = P.dphi_p; o.d.epi.raw_C = dq;
            if (build(&P.ad_phi, o, &P.flops_bwd, &P.launches_bwd)) return nullptr;
        }
        {   // dg_p = P^T dO
            OpB o(P.PT, b, Hk, Hk, Nq, 0, Nq, P.dOT, dv, 1, EPI_FWD);
            o.d.B_batch = b; o.d.epi.raw = P.dg_p; o.d.epi.raw_C = dv;
            if (build(&P.ad_g, o, &P.flops_bwd, &P.launches_bwd)) return nullptr;
        }
        {   // dx = dqkv W_qkv + dh (identity path of x + gamma*o)
            OpB o(P.dqkv, b, H, H, nq, 0, nq, attn.wqkv_t, C, 1, EPI_BWD);
            ConvGemmParams& e = o.d.epi;
            e.addin = dh_attn_out; e.addin_C = C; e.addin_climit = C; e.addin_pool = 0;
            e.dx = dh_attn_in; e.dx_C = C;
            if (build(&P.ad_qkv, o, &P.flops_bwd, &P.launches_bwd)) return nullptr;
        }
    }
    // ---- rgb
    {
        const BigGANPlan::BB& Bl = P.bb[nL - 1];
        OpB o(Bl.out_act, b, R, R, C_last, 0, C_last, wrgb, 27, 1, EPI_FWD);
        o.d.epi.img_nchw = P.rgbT;  // planar fp32 [b, 27, R, R]; k_rgb_gather adds bias, sums the taps, tanh
        o.d.epi.img_linear = 1;
        if (build(&P.f_rgb, o, &P.flops_fwd, &P.launches_fwd)) return nullptr;
        P.launches_fwd += 1;
    }
    {
        const BigGANPlan::BB& Bl = P.bb[nL - 1];
        OpB o(P.col_rgb, b, R, R, 64, 0, 64, wrgb_t, C_last, 1, EPI_BWD);
        ConvGemmParams& e = o.d.epi;
        e.saved = Bl.out_act; e.saved_C = C_last;
        e.aff_a = P.a + bns[final_bn].off; e.aff_stride = C_all;
        e.dx = P.dhA; e.dx_C = C_last;  // block nL-1 reads dh_out from dhA
        if (build(&P.d_rgb, o, &P.flops_bwd, &P.launches_bwd)) return nullptr;
        P.launches_bwd += 1;  // im2col
    }
    // ---- serpentine tile order: the k-th tensor-core launch of a pass walks its tiles backwards when k is odd, so that
    // every layer starts on the tiles its producer wrote last (still in L2) — same results, the BN-gradient partial
    // slots are indexed by tile, not by arrival order
    if (get_option("serpentine") != 0) {
        std::vector<ConvOp*> fs, bs;
        for (int i = 0; i < nL; ++i) {
            if (attn.C && i == cfg.attention_pos) {
                fs.push_back(&P.a_qkv);
                if (P.attn_fused) { fs.push_back(&P.a_s1); fs.push_back(&P.a_s2); } else fs.push_back(&P.a_s);
                fs.push_back(&P.a_o);
                fs.push_back(&P.a_out);
            }
            for (int k = 0; k < 4; ++k) fs.push_back(&P.bb[i].f[k]);
        }
        fs.push_back(&P.f_rgb);
        bs.push_back(&P.d_rgb);
        for (int i = nL - 1; i >= 0; --i) {
            for (int k = 3; k >= 0; --k) bs.push_back(&P.bb[i].d[k]);
            if (attn.C && i == cfg.attention_pos) {
                bs.push_back(&P.ad_out);
                bs.push_back(P.attn_fused ? &P.ad_pf : &P.ad_p);
                bs.push_back(&P.ad_theta); bs.push_back(&P.ad_phi); bs.push_back(&P.ad_g); bs.push_back(&P.ad_qkv);
            }
        }
        for (size_t k = 0; k < fs.size(); ++k) fs[k]->p.tile_reverse = (int)(k & 1);
        for (size_t k = 0; k < bs.size(); ++k) bs[k]->p.tile_reverse = (int)(k & 1);
    }
    P.launches_fwd += 4;  // concat, cond_affine, uncond_affine, gen_z
    P.launches_bwd += 6;  // memsets + finalize + dcond x2 + convert + split
    BigGANPlan* raw = pp.get();
    plans[b] = pp;
    return raw;
}

// ----------------------------------------------------------------------------- forward
int BigGAN::forward(int b, const float* z, const float* c, float* img, cudaStream_t st) {
    if (!finalized) { set_error("biggan: forward before finalize"); return -1; }
    BigGANPlan* Pp = plan(b);
    if (!Pp) return -1;
    BigGANPlan& P = *Pp;
    last_plan = Pp;
    k_concat_cond(z, c, P.cond, b, cfg.z_dim, cfg.class_embed_dim, st);
    k_cond_affine(P.cond, WT_as, bias_as, P.a, P.s, b, cdim, C_cond, C_all, st);
    k_uncond_affine(unc_weight, unc_bias, mean, inv_std, P.a, P.s, b, C_cond, C_last, C_all, st);
    const int bn00 = blocks[0].bn[0];
    // gen_z output is already NHWC: view(b, 4, 4, C0)
    k_gen_z(P.cond, genz_WT, genz_b, P.a + bns[bn00].off, P.s + bns[bn00].off, C_all, P.bb[0].in_raw, P.bb[0].in_act,
            b, cdim, genz_J, C0, st);
    const int nL = (int)blocks.size();
    for (int i = 0; i < nL; ++i) {
        if (attn.C && i == cfg.attention_pos) {
            const int H = attn.H, dq = attn.dq, dv = attn.dv, nq = 2 * dq + dv;
            if (conv_op_launch(P.a_qkv, st)) return -1;
            k_maxpool2_fwd(P.qkv, nq, dq, dq, P.phi_p, P.phiT, P.idx_phi, b, H, H, st);
            k_maxpool2_fwd(P.qkv, nq, 2 * dq, dv, P.g_p, P.gT, P.idx_g, b, H, H, st);
            if (P.attn_fused) {
                if (conv_op_launch(P.a_s1, st)) return -1;
                if (conv_op_launch(P.a_s2, st)) return -1;
            } else {
                if (conv_op_launch(P.a_s, st)) return -1;
                k_softmax_fwd(P.S, P.P, (long)b * H * H, H * H / 4, st);
            }
            if (conv_op_launch(P.a_o, st)) return -1;
            if (conv_op_launch(P.a_out, st)) return -1;
        }
        for (int k = 0; k < 4; ++k)
            if (conv_op_launch(P.bb[i].f[k], st)) return -1;
    }
    if (conv_op_launch(P.f_rgb, st)) return -1;
    k_rgb_gather(P.rgbT, brgb, img ? img : P.img, b, H_out, H_out, st);
    if (img && img != P.img) {
        // keep a private copy for the backward pass (tanh')
        P2L_CUDA_CHECK(cudaMemcpyAsync(P.img, img, (size_t)b * 3 * H_out * H_out * sizeof(float),
                                       cudaMemcpyDeviceToDevice, st));
    }
    P.forward_done = true;
    return 0;
}

// ----------------------------------------------------------------------------- backward
int BigGAN::backward(int b, const float* dimg, float* dz, float* dc, cudaStream_t st, float scale,
                     const float* row_scale) {
    auto it = plans.find(b);
    if (it == plans.end() || !it->second->forward_done) {
        set_error("biggan: backward(b=%d) without a matching forward", b);
        return -1;
    }
    BigGANPlan& P = *it->second;
    const int nL = (int)blocks.size();
    const int R = H_out;
    // image -> last block output
    k_im2col_rgb_bwd(dimg, P.img, P.col_rgb, b, R, R, 64, grad_scale(), st);  // 16-bit gradients carry grad_scale() from here ...
    if (conv_op_launch(P.d_rgb, st)) return -1;
    for (int i = nL - 1; i >= 0; --i) {
        const Block& bl = blocks[i];
        BigGANPlan::BB& B = P.bb[i];
        if (conv_op_launch(B.d[3], st)) return -1;
        if (conv_op_launch(B.d[2], st)) return -1;
        if (conv_op_launch(B.d[1], st)) return -1;
        if (bl.up) {
            const BN& bn1 = bns[bl.bn[1]];
            k_pool_bnrelu_bwd(P.g3, B.t1_lo, P.a + bn1.off, C_all, B.sp[1], P.g1, b, bl.Hin, bl.Hin, bl.mid, st);
            act_t* dh_out = ((nL - 1 - i) % 2 == 0) ? P.dhA : P.dhB;
            k_pool2x2_sum(dh_out, bl.out, P.dh_pool, b, bl.Hin, bl.Hin, bl.out, st);
        }
        if (conv_op_launch(B.d[0], st)) return -1;
        if (attn.C && i == cfg.attention_pos) {
            const int H = attn.H, dq = attn.dq, dv = attn.dv, nq = 2 * dq + dv;
            const int Nq = H * H, Nk = Nq / 4;
            if (conv_op_launch(P.ad_out, st)) return -1;              // dO
            if (P.attn_fused) {
                k_rowdot(P.dO, P.O, P.Drow, (long)b * Nq, dv, st);    // rowsum(dP o P) = dO . O
                if (conv_op_launch(P.ad_pf, st)) return -1;           // dS
            } else {
                if (conv_op_launch(P.ad_p, st)) return -1;            // dP -> S
                k_softmax_bwd(P.P, P.S, P.dS, (long)b * Nq, Nk, st);  // dS
            }
            if (conv_op_launch(P.ad_theta, st)) return -1;            // dtheta -> dqkv[:, :dq]
            if (!P.attn_emit_t) {
                k_transpose(P.dS, Nk, 0, P.dST, b, Nq, Nk, st);
                k_transpose(P.qkv, nq, 0, P.thetaT, b, Nq, dq, st);
            }
            if (conv_op_launch(P.ad_phi, st)) return -1;              // dphi_p
            if (!P.attn_emit_t) {
                k_transpose(P.P, Nk, 0, P.PT, b, Nq, Nk, st);
                k_transpose(P.dO, dv, 0, P.dOT, b, Nq, dv, st);
            }
            if (conv_op_launch(P.ad_g, st)) return -1;                // dg_p
            k_maxpool2_bwd(P.dphi_p, P.idx_phi, P.dqkv, nq, dq, dq, b, H, H, st);
            k_maxpool2_bwd(P.dg_p, P.idx_g, P.dqkv, nq, 2 * dq, dv, b, H, H, st);
            if (conv_op_launch(P.ad_qkv, st)) return -1;              // -> dh of block ap-1
        }
    }
    // BN-affine gradients -> d cond: fixed-order sums of the per-tile partials, then the finalisation
    k_stat_reduce(P.segs, P.nsegs, P.S0, P.S1, C_all, b, st);
    k_bn_grad_finalize(P.S0, P.S1, P.a, P.s, mean, inv_std, P.G, b, C_cond, C_all, st);
    const int nb_g = k_dcond_blocks(2 * C_cond), nb_z = k_dcond_blocks(genz_J);
    k_dcond_partial(P.G, 2 * C_cond, Wcat, P.dcond, b, 2 * C_cond, cdim, st);
    k_dcond_partial(P.dh0, genz_J, genz_W, P.dcond + (size_t)nb_g * b * cdim, b, genz_J, cdim, st);
    k_dcond_reduce_split(P.dcond, nb_g + nb_z, dz, dc, b, cfg.z_dim, cfg.class_embed_dim, scale / grad_scale(), row_scale, st);  // ... to here
    return 0;
}

const float* BigGAN::last_image(int b) {
    auto it = plans.find(b);
    return it == plans.end() ? nullptr : it->second->img;
}
size_t BigGAN::device_bytes() {
    size_t t = weights.total;
    for (auto& kv : plans) t += kv.second->ar.total;
    return t;
}
double BigGAN::flops(int b, int backward) {
    BigGANPlan* P = plan(b);
    if (!P) return 0;
    return backward ? P->flops_bwd : P->flops_fwd;
}

}  // namespace p2l
