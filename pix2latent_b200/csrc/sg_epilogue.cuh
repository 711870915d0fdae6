// StyleGAN2 modulated-convolution epilogues of the tcgen05 implicit-GEMM kernels (conv_gemm.cuh). Included by
// conv_gemm.cuh; a separate, lean epilogue so that the BigGAN kernels' hot loop does not grow (FLAVOR_SG instantiations).
//
// rosinality's ModulatedConv2d + NoiseInjection + FusedLeakyReLU (oracle/stylegan2.py, reached through
// /root/reference pix2latent/model/stylegan2.py:110-119) as an input-channel scale before and an output-channel scale
// after a SHARED-weight convolution:
//
//   forward   pre = dm[n,c] * acc + nw * noise[n,pix] + bias[c] ;  x = lrelu(pre) * sqrt2          -> raw  (x_l)
//             A_next = s_next[n,c] * x                                                              -> act  (operand of layer l+1)
//             up-sampling layers run as FOUR phase convolutions on the low-resolution grid (conv_transpose(stride 2)
//             composed with the 4x4 FIR blur = one 3x3 filter per output phase): the GEMM's columns are (phase, channel)
//             and the stores go depth-to-space (d2s_C = channels per phase).
//
//   backward  (epilogue of layer l's dgrad, acc = dL/dA_l; `saved` = x_{l-1}, the UNmodulated input of layer l)
//             ds_l[n,c]    = sum_pix acc * x_{l-1}                      (gradient of the modulation, stat 0)
//             dx           = s_l[n,c] * acc + addin                      (addin: gradient through the ToRGB branch)
//             g            = dx * sqrt2 * lrelu'(x_{l-1})
//             u * dm       = lrelu^-1(x_{l-1} / sqrt2) - nw * noise - bias   (the conv output before demodulation is u; recomputed)
//             (ddm * dm)_{l-1}[n,c] = sum_pix g * u * dm                  (gradient of the demodulation TIMES dm: the per-column
//                                                                          division is left to the consumer, k_demod_bwd; stat 1)
//             G_{l-1}      = dm_{l-1}[n,c] * g                           -> dx (operand of layer l-1's dgrad; space-to-depth
//                                                                          when layer l-1 is an up-sampling layer)
// i.e. one epilogue = the reference's modulate-backward of layer l + noise/bias/activation/demodulation-backward of layer
// l-1; no element-wise pass touches HBM in between. The two per-(sample, channel) sums use the deterministic per-tile
// partial slots of the BN-gradient sums (ConvGemmParams::statp).
#pragma once

namespace p2l {

constexpr float kSgSqrt2 = 1.41421356237309515f;

template <int BN, int MODE, int NG>
__device__ __forceinline__ void epilogue_loop_sg(const ConvGemmParams& p, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                                 uint32_t tmem_base, int total_tiles, int warp, int lane, float* ctab) {
    constexpr int CH = 32;
    static_assert(BN % CH == 0, "StyleGAN2 epilogues run on 32-column chunks");
    float* stat_red = ctab + 6 * BN;
    int stat_it = 0;
    const bool use_tab = (p.nb == 1);
    const int grp = (NG == 2) ? ((warp - 2) >> 2) : 0;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int wi = row % p.tw;
    const int hi = (row / p.tw) % p.th;
    const int ni = row / (p.tw * p.th);
    const int rows_per_img = p.tw * p.th;
    const float nw = p.sg_nw ? __ldg(p.sg_nw) : 0.f;
    const int Cc = p.d2s_C > 0 ? p.d2s_C : p.Cout;   // channels per phase: table / coefficient index = column % Cc
    int it = grp;
    for (int wt = blockIdx.x + grp * gridDim.x; wt < total_tiles; wt += NG * gridDim.x, it += NG) {
        const int tile = tile_of(p, wt, total_tiles);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
        const int twi = m_tile % p.tiles_w;
        const int thi = (m_tile / p.tiles_w) % p.tiles_h;
        const int tni = m_tile / (p.tiles_w * p.tiles_h);
        const int w = twi * p.tw + wi, h = thi * p.th + hi, n = tni * p.nb + ni;
        const bool valid = (w < p.W) && (h < p.H) && (n < p.NI);
        const int nn = valid ? n : 0;
        const long pix = (static_cast<long>(n) * p.H + h) * p.W + w;
        if constexpr (MODE == EPI_BWD) {
            if (p.saved && valid) prefetch_l2_row<BN * 2>(p.saved + pix * p.saved_C + n_tile * BN);   // see epilogue_loop_direct
        }
        // per-tile coefficient tables: [0] bias, [1] aff_a (FWD: next layer's modulation; BWD: this layer's), [2] dm
        const uint32_t tab = smem_u32(ctab) + (it & 1) * 3 * BN * 4;
        if (use_tab) {
            if constexpr (NG == 2) bar_epilogue(grp);
            const int et = ((warp - 2) & 3) * 32 + lane;
            const int nt = min(tni * p.nb, p.NI - 1);
            for (int j = et; j < BN; j += 128) {
                const int col = n_tile * BN + j;
                const bool in = col < p.Cout;
                const int ch = col % Cc;
                const float* bp = (MODE == EPI_FWD) ? p.bias : p.sg_bias;
                sts32f(tab + j * 4, (bp && in) ? __ldg(bp + ch) : 0.f);
                sts32f(tab + (BN + j) * 4, (p.aff_a && in) ? __ldg(p.aff_a + static_cast<long>(nt) * p.aff_stride + ch) : 1.f);
                sts32f(tab + (2 * BN + j) * 4, (p.sg_dm && in) ? __ldg(p.sg_dm + static_cast<long>(nt) * p.sg_ld + ch) : 1.f);
            }
            bar_epilogue(grp);
        }
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
#pragma unroll 1
        for (int c = 0; c < BN; c += CH) {
            const int cbase = n_tile * BN + c;
            if (cbase >= p.Cout) break;
            float v[CH];
            {
                uint32_t u[CH];
                tmem_ld32(t_addr + c, u);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(u[j]);
            }
            const int cch = cbase % Cc;                 // channel of column cbase
            // coefficient k (0 bias, 1 aff_a, 2 dm) of columns j .. j+3 of this chunk: from the per-tile table, or (tiles
            // that span several small images) from global memory for this thread's image
            const float* bp = (MODE == EPI_FWD) ? p.bias : p.sg_bias;
            auto coef4 = [&](int k, int j) -> float4 {
                if (use_tab) return lds128f(tab + (k * BN + c + j) * 4);
                if (k == 0) return bp ? __ldg(reinterpret_cast<const float4*>(bp + cch + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (k == 1) return p.aff_a ? __ldg(reinterpret_cast<const float4*>(p.aff_a + static_cast<long>(nn) * p.aff_stride + cch + j)) : make_float4(1.f, 1.f, 1.f, 1.f);
                return p.sg_dm ? __ldg(reinterpret_cast<const float4*>(p.sg_dm + static_cast<long>(nn) * p.sg_ld + cch + j)) : make_float4(1.f, 1.f, 1.f, 1.f);
            };
            if constexpr (MODE == EPI_FWD) {
                // output pixel: the conv grid's, or (depth-to-space) phase (py, px) of the 2x grid
                int oy = h, ox = w, oH = p.H, oW = p.W;
                if (p.d2s_C > 0) {
                    const int ph = cbase / p.d2s_C;
                    oy = 2 * h + (ph >> 1); ox = 2 * w + (ph & 1); oH = 2 * p.H; oW = 2 * p.W;
                }
                const long opix = (static_cast<long>(n) * oH + oy) * oW + ox;
                const float nz = (p.sg_noise && valid) ? nw * __ldg(p.sg_noise + opix) : 0.f;
#pragma unroll
                for (int q = 0; q < CH / 4; ++q) {
                    const float4 b4 = coef4(0, q * 4), d4 = coef4(2, q * 4);
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float pre = fmaf(dd[e], v[q * 4 + e], bb[e] + nz);
                        v[q * 4 + e] = (pre > 0.f ? pre : 0.2f * pre) * kSgSqrt2;
                    }
                }
                if (valid) {
                    if (p.raw) row_store<CH>(p.raw + opix * p.raw_C + cch, v, wide_ok(p.raw, p.raw_C * 2));
                    if (p.act) {
#pragma unroll
                        for (int q = 0; q < CH / 4; ++q) {
                            const float4 a4 = coef4(1, q * 4);
                            v[q * 4] *= a4.x; v[q * 4 + 1] *= a4.y; v[q * 4 + 2] *= a4.z; v[q * 4 + 3] *= a4.w;
                        }
                        row_store<CH>(p.act + opix * p.act_C + cch, v, wide_ok(p.act, p.act_C * 2));
                    }
                }
            } else {
                float y[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j) y[j] = 0.f;
                if (valid) row_load_add<CH>(p.saved + pix * p.saved_C + cbase, y, wide_ok(p.saved, p.saved_C * 2));
                else {
#pragma unroll
                    for (int j = 0; j < CH; ++j) v[j] = 0.f;
                }
                float s0 = 0.f, s1 = 0.f, h0[2] = {0.f, 0.f}, h1[2] = {0.f, 0.f};   // column sums (32-row / 16-row groups)
                const bool rows32 = rows_per_img >= 32;
                {   // ds_l partial: sum_pix acc * x
                    float t[CH];
#pragma unroll
                    for (int j = 0; j < CH; ++j) t[j] = v[j] * y[j];
                    if (p.statp) {
                        if (rows32) s0 = colsum_group<32>(t, lane);
                        else {
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                float a[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j) a[j] = t[half * 16 + j];
                                h0[half] = colsum_group<16>(a, lane);
                            }
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < CH / 4; ++q) {   // dx = s * acc
                    const float4 a4 = coef4(1, q * 4);
                    v[q * 4] *= a4.x; v[q * 4 + 1] *= a4.y; v[q * 4 + 2] *= a4.z; v[q * 4 + 3] *= a4.w;
                }
                if (p.addin && valid) row_load_add<CH>(p.addin + pix * p.addin_C + cbase, v, wide_ok(p.addin, p.addin_C * 2));
                if (p.dx2 && valid) row_store<CH>(p.dx2 + pix * p.dx_C + cbase, v, wide_ok(p.dx2, p.dx_C * 2));
                const float nz = (p.sg_noise && valid) ? nw * __ldg(p.sg_noise + pix) : 0.f;
                {   // through lrelu * sqrt2 and the demodulation of the layer that produced `saved`
                    float t[CH];
#pragma unroll
                    for (int q = 0; q < CH / 4; ++q) {
                        const float4 b4 = coef4(0, q * 4), d4 = coef4(2, q * 4);
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = q * 4 + e;
                            const bool pos = y[j] > 0.f;
                            const float g = v[j] * (pos ? kSgSqrt2 : 0.2f * kSgSqrt2);
                            const float pre = y[j] * (pos ? (1.f / kSgSqrt2) : (1.f / (0.2f * kSgSqrt2)));
                            t[j] = valid ? g * (pre - nz - bb[e]) : 0.f;   // = g * u * dm: the division by dm[n,c] waits for the column sum
                            v[j] = dd[e] * g;
                        }
                    }
                    if (p.statp) {
                        if (rows32) s1 = colsum_group<32>(t, lane);
                        else {
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                float a[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j) a[j] = t[half * 16 + j];
                                h1[half] = colsum_group<16>(a, lane);
                            }
                        }
                    }
                }
                if (p.statp) {
                    // per-(sample, channel) sums over the pixels: slots and order as in epilogue_loop_direct
                    if (rows32) {
                        if (p.nb == 1) {
                            const uint32_t red = smem_u32(stat_red) + ((grp * 2 + (stat_it & 1)) * 256) * 4;
                            ++stat_it;
                            sts32f(red + (quad * 64 + lane) * 4, s0);
                            sts32f(red + (quad * 64 + 32 + lane) * 4, s1);
                            bar_epilogue(grp);
                            if (quad == 0 && tni < p.NI && cbase + lane < p.Cout) {
                                s0 = (lds32f(red + lane * 4) + lds32f(red + (64 + lane) * 4)) + (lds32f(red + (128 + lane) * 4) + lds32f(red + (192 + lane) * 4));
                                s1 = (lds32f(red + (32 + lane) * 4) + lds32f(red + (96 + lane) * 4)) + (lds32f(red + (160 + lane) * 4) + lds32f(red + (224 + lane) * 4));
                                const int part = thi * p.tiles_w + twi;
                                float* dst = p.statp + (static_cast<long>(tni) * p.statp_parts + part) * 2 * p.statp_C + cbase + lane;
                                dst[0] = s0;
                                dst[p.statp_C] = s1;
                            }
                        } else {
                            const int ns = tni * p.nb + (quad * 32) / rows_per_img;
                            const int part = ((quad * 32) % rows_per_img) >> 5;
                            if (ns < p.NI && cbase + lane < p.Cout) {
                                float* dst = p.statp + (static_cast<long>(ns) * p.statp_parts + part) * 2 * p.statp_C + cbase + lane;
                                dst[0] = s0;
                                dst[p.statp_C] = s1;
                            }
                        }
                    } else {
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const int ns = tni * p.nb + (quad * 32 + (lane & 16)) / rows_per_img;
                            const int ch = cbase + half * 16 + (lane & 15);
                            if (ns < p.NI && ch < p.Cout) {
                                float* dst = p.statp + static_cast<long>(ns) * p.statp_parts * 2 * p.statp_C + ch;
                                dst[0] = h0[half];
                                dst[p.statp_C] = h1[half];
                            }
                        }
                    }
                }
                if (p.dx && valid) {
                    if (p.s2d) {
                        // space-to-depth: [NI, H/2, W/2, 4 * dx_C], channel block (h & 1) * 2 + (w & 1)
                        const long lp = (static_cast<long>(n) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
                        row_store<CH>(p.dx + (lp * 4 + ((h & 1) * 2 + (w & 1))) * p.dx_C + cbase, v, wide_ok(p.dx, p.dx_C * 2));
                    } else {
                        row_store<CH>(p.dx + pix * p.dx_C + cbase, v, wide_ok(p.dx, p.dx_C * 2));
                    }
                }
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
}

}  // namespace p2l
