// Implicit-GEMM NHWC convolution / batched GEMM on tcgen05 tensor cores (sm_100a).
//
//   D[pixel, cout] = sum_{tap, cin} A[pixel + tap, cin] * B[cout, tap*Cin + cin]
//
// * A (activations, bf16 NHWC) is fetched by TMA as a 4-D box {64 ch, tw, th, nb} per filter
//   tap; out-of-range coordinates are zero-filled by the TMA unit, which is the conv padding.
// * B (weights, bf16 [Cout, taps*Cin], K contiguous) is fetched by TMA as {64, BN}.
// * Both land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma reads.
// * One elected thread issues tcgen05.mma (M=128, N=BN, K=16) into a TMEM accumulator; two
//   accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
// * Persistent CTAs (one per SM) walk tiles round-robin.
// * Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM owner), 2..5 = epilogue.
//
// The epilogue is where the reference's element-wise layers live:
//   FWD: v = alpha*acc + bias (+ residual skip) ; raw = v ; act = relu(a[b,c]*v + s[b,c])
//        (conditional-BN affine of the *next* layer, pytorch_pretrained_biggan BigGANBatchNorm),
//        optional nearest-x2 replicated write, optional tanh + NCHW fp32 image write.
//   BWD: g = acc ; dpre = g * [saved_act > 0] ; per-(sample,channel) sums of dpre and
//        dpre*saved_act (the BN-affine gradients) ; dx = a[b,c]*dpre (+ skip gradient).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace p2l {

constexpr int kBM = 128;  // pixels per tile
constexpr int kBK = 64;   // bf16 per K chunk = one 128-byte swizzle row
constexpr int kGemmThreads = 192;
constexpr int kATileBytes = kBM * kBK * 2;

enum { EPI_FWD = 0, EPI_BWD = 1 };

struct ConvGemmParams {
    // ---- M: output pixel grid [NI, H, W], tile box (nb, th, tw), tw*th*nb == 128
    int NI, H, W;
    int tw, th, nb;
    int tiles_w, tiles_h, tiles_n;
    // ---- N
    int Cout;     // logical output channels
    int n_tiles;  // ceil(Cout / BN)
    // ---- K
    int taps_h, taps_w, pad_h, pad_w;
    int cin_chunks;  // Cin / 64
    int a_c0;        // channel offset into the A tensor
    int b_batched;   // B tensor map's 3rd coordinate = image index
    // ---- epilogue, forward
    float alpha;             // scale on the accumulator
    const float* alpha_ptr;  // optional device scalar multiplied into alpha (attention gamma); both modes
    const float* bias;       // [Cout] or null
    const __nv_bfloat16* resid;  // skip input, NHWC [NI, H>>resid_shift, W>>resid_shift, resid_C]
    int resid_C, resid_shift;
    __nv_bfloat16* raw;  // raw (pre-affine) output, NHWC, channel stride raw_C
    int raw_C;
    float* raw_f32;  // same, fp32 (attention logits)
    int raw_f32_C;
    const float* aff_a;  // per-sample affine [NI, aff_stride] (pointer pre-offset to this layer)
    const float* aff_s;
    int aff_stride;
    int relu;
    __nv_bfloat16* act;  // activated output, NHWC, channel stride act_C
    int act_C;
    int act_up;             // write act to the 2x nearest-upsampled grid [NI, 2H, 2W, act_C]
    __nv_bfloat16* act_lo;  // with act_up: also keep the low-res copy (needed by backward)
    float* img_nchw;        // tanh(v) for c < Cout written as fp32 NCHW [NI, Cout, H, W]
    // ---- epilogue, backward
    const __nv_bfloat16* saved;  // forward activation of the layer being differentiated
    int saved_C;
    float* stat0;  // += sum_pix dpre          [NI, stat_stride]
    float* stat1;  // += sum_pix dpre * saved  [NI, stat_stride]
    int stat_stride;
    const __nv_bfloat16* addin;  // gradient arriving through the skip connection
    int addin_C, addin_climit, addin_pool;  // pool: sum the 2x2 block of a [NI,2H,2W,addin_C] map
    __nv_bfloat16* dx;
    int dx_C;
    float* dx_f32;  // optional fp32 copy (used for the latent-side tensors)
    int dx_f32_C;
};

template <int BN>
struct GemmCfg {
    static constexpr int kBTileBytes = BN * kBK * 2;
    static constexpr int kStageBytes = kATileBytes + kBTileBytes;
    static constexpr int kMaxStages = (200 * 1024) / kStageBytes;
    static constexpr int kStages = kMaxStages > 8 ? 8 : kMaxStages;
    static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N constraint for M=128");
};

// Sum each of G columns over the G lanes of a lane group; lane l ends up with column (l % G).
template <int G>
__device__ __forceinline__ float colsum_group(float (&v)[G], int lane) {
#pragma unroll
    for (int h = G / 2; h >= 1; h >>= 1) {
        const bool upper = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = upper ? v[i] : v[i + h];
            const float keep = upper ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    return v[0];
}

template <int BN, int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvGemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int S = Cfg::kStages;
    constexpr int CH = (BN >= 32) ? 32 : 16;  // epilogue column chunk

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + S;
    uint64_t* tfull_bar = bars + 2 * S;
    uint64_t* tempty_bar = bars + 2 * S + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tfull_bar[0], 1);
        mbar_init(&tfull_bar[1], 1);
        mbar_init(&tempty_bar[0], 4);
        mbar_init(&tempty_bar[1], 4);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int total_tiles = m_tiles * p.n_tiles;
    const int k_blocks = p.taps_h * p.taps_w * p.cin_chunks;
    const int Cin = p.cin_chunks * kBK;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
                const int twi = m_tile % p.tiles_w;
                const int thi = (m_tile / p.tiles_w) % p.tiles_h;
                const int tni = m_tile / (p.tiles_w * p.tiles_h);
                const int w0 = twi * p.tw, h0 = thi * p.th, n0 = tni * p.nb;
                for (int r = 0; r < p.taps_h; ++r) {
                    for (int s = 0; s < p.taps_w; ++s) {
                        const int kbase = (r * p.taps_w + s) * Cin;
                        for (int cc = 0; cc < p.cin_chunks; ++cc) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            uint8_t* sA = smem + stage * Cfg::kStageBytes;
                            uint8_t* sB = sA + kATileBytes;
                            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                            tma_load_4d(sA, &tmA, &full_bar[stage], p.a_c0 + cc * kBK,
                                        w0 + s - p.pad_w, h0 + r - p.pad_h, n0);
                            tma_load_3d(sB, &tmB, &full_bar[stage], kbase + cc * kBK, n_tile * BN,
                                        p.b_batched ? n0 : 0);
                            if (++stage == S) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint64_t adesc = umma_desc_k128(sA);
                    const uint64_t bdesc = umma_desc_k128(sA + kATileBytes);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        // advance 16 bf16 = 32 B along K inside the swizzle row: +2 in 16-B units
                        umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull_bar[as]);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int quad = warp & 3;  // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;
        const int wi = row % p.tw;
        const int hi = (row / p.tw) % p.th;
        const int ni = row / (p.tw * p.th);
        const int rows_per_img = p.tw * p.th;
        float alpha = p.alpha;
        if (p.alpha_ptr) alpha *= __ldg(p.alpha_ptr);
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
            const int twi = m_tile % p.tiles_w;
            const int thi = (m_tile / p.tiles_w) % p.tiles_h;
            const int tni = m_tile / (p.tiles_w * p.tiles_h);
            const int w = twi * p.tw + wi, h = thi * p.th + hi, n = tni * p.nb + ni;
            const bool valid = (w < p.W) && (h < p.H) && (n < p.NI);
            const long pix = (static_cast<long>(n) * p.H + h) * p.W + w;

            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;

#pragma unroll 1
            for (int c = 0; c < BN; c += CH) {
                const int cbase = n_tile * BN + c;
                if (cbase >= p.Cout) break;  // warp-uniform
                float v[CH];
                {
                    uint32_t u[CH];
                    if constexpr (CH == 32) tmem_ld32(t_addr + c, u);
                    else tmem_ld16(t_addr + c, u);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(u[j]);
                }
                const bool full_chunk = (cbase + CH <= p.Cout);

                if constexpr (MODE == EPI_FWD) {
                    // ---- v = alpha*acc + bias (+ skip)
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        float b = 0.f;
                        if (p.bias && (full_chunk || cbase + j < p.Cout)) b = __ldg(p.bias + cbase + j);
                        v[j] = alpha * v[j] + b;
                    }
                    if (p.resid && valid) {
                        const int Hs = p.H >> p.resid_shift, Ws = p.W >> p.resid_shift;
                        const long rp = (static_cast<long>(n) * Hs + (h >> p.resid_shift)) * Ws + (w >> p.resid_shift);
                        const uint4* src = reinterpret_cast<const uint4*>(p.resid + rp * p.resid_C + cbase);
#pragma unroll
                        for (int q = 0; q < CH / 8; ++q) {
                            const uint4 t = __ldg(src + q);
                            v[q * 8 + 0] += bf16_lo(t.x); v[q * 8 + 1] += bf16_hi(t.x);
                            v[q * 8 + 2] += bf16_lo(t.y); v[q * 8 + 3] += bf16_hi(t.y);
                            v[q * 8 + 4] += bf16_lo(t.z); v[q * 8 + 5] += bf16_hi(t.z);
                            v[q * 8 + 6] += bf16_lo(t.w); v[q * 8 + 7] += bf16_hi(t.w);
                        }
                    }
                    if (p.img_nchw) {
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < CH; ++j) {
                                if (cbase + j < p.Cout) {
                                    p.img_nchw[((static_cast<long>(n) * p.Cout + cbase + j) * p.H + h) * p.W + w] = tanhf(v[j]);
                                }
                            }
                        }
                    }
                    if (p.raw_f32 && valid) {
                        float4* dst = reinterpret_cast<float4*>(p.raw_f32 + pix * p.raw_f32_C + cbase);
#pragma unroll
                        for (int q = 0; q < CH / 4; ++q) dst[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                    }
                    if (p.raw && valid) {
                        uint4* dst = reinterpret_cast<uint4*>(p.raw + pix * p.raw_C + cbase);
#pragma unroll
                        for (int q = 0; q < CH / 8; ++q) {
                            dst[q] = make_uint4(pack_bf16(v[q * 8], v[q * 8 + 1]), pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                                                pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
                        }
                    }
                    if (p.act) {
                        if (p.aff_a) {
                            const int nn = valid ? n : 0;
                            const float* pa = p.aff_a + static_cast<long>(nn) * p.aff_stride + cbase;
                            const float* ps = p.aff_s + static_cast<long>(nn) * p.aff_stride + cbase;
#pragma unroll
                            for (int q = 0; q < CH / 4; ++q) {
                                const float4 a4 = __ldg(reinterpret_cast<const float4*>(pa) + q);
                                const float4 s4 = __ldg(reinterpret_cast<const float4*>(ps) + q);
                                v[q * 4 + 0] = fmaf(a4.x, v[q * 4 + 0], s4.x);
                                v[q * 4 + 1] = fmaf(a4.y, v[q * 4 + 1], s4.y);
                                v[q * 4 + 2] = fmaf(a4.z, v[q * 4 + 2], s4.z);
                                v[q * 4 + 3] = fmaf(a4.w, v[q * 4 + 3], s4.w);
                            }
                        }
                        if (p.relu) {
#pragma unroll
                            for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.f);
                        }
                        if (valid) {
                            uint4 o[CH / 8];
#pragma unroll
                            for (int q = 0; q < CH / 8; ++q) {
                                o[q] = make_uint4(pack_bf16(v[q * 8], v[q * 8 + 1]), pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                                                  pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
                            }
                            if (!p.act_up) {
                                uint4* dst = reinterpret_cast<uint4*>(p.act + pix * p.act_C + cbase);
#pragma unroll
                                for (int q = 0; q < CH / 8; ++q) dst[q] = o[q];
                            } else {
                                const int W2 = p.W * 2, H2 = p.H * 2;
#pragma unroll
                                for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
                                    for (int dxx = 0; dxx < 2; ++dxx) {
                                        const long hp = (static_cast<long>(n) * H2 + 2 * h + dy) * W2 + 2 * w + dxx;
                                        uint4* dst = reinterpret_cast<uint4*>(p.act + hp * p.act_C + cbase);
#pragma unroll
                                        for (int q = 0; q < CH / 8; ++q) dst[q] = o[q];
                                    }
                                }
                                if (p.act_lo) {
                                    uint4* dst = reinterpret_cast<uint4*>(p.act_lo + pix * p.act_C + cbase);
#pragma unroll
                                    for (int q = 0; q < CH / 8; ++q) dst[q] = o[q];
                                }
                            }
                        }
                    }
                } else {
                    // ---------------------------------------------------------- backward
                    float y[CH];
#pragma unroll
                    for (int j = 0; j < CH; ++j) { y[j] = 0.f; v[j] *= alpha; }
                    if (p.saved) {
                        if (valid) {
                            const uint4* src = reinterpret_cast<const uint4*>(p.saved + pix * p.saved_C + cbase);
#pragma unroll
                            for (int q = 0; q < CH / 8; ++q) {
                                const uint4 t = __ldg(src + q);
                                y[q * 8 + 0] = bf16_lo(t.x); y[q * 8 + 1] = bf16_hi(t.x);
                                y[q * 8 + 2] = bf16_lo(t.y); y[q * 8 + 3] = bf16_hi(t.y);
                                y[q * 8 + 4] = bf16_lo(t.z); y[q * 8 + 5] = bf16_hi(t.z);
                                y[q * 8 + 6] = bf16_lo(t.w); y[q * 8 + 7] = bf16_hi(t.w);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < CH; ++j) y[j] = 0.f;
                        }
#pragma unroll
                        for (int j = 0; j < CH; ++j) v[j] = (y[j] > 0.f) ? v[j] : 0.f;
                    } else if (!valid) {
#pragma unroll
                        for (int j = 0; j < CH; ++j) v[j] = 0.f;
                    }
                    if (p.stat0) {
                        // BN-affine gradients: reduce over the pixels (lanes) of one image.
                        if (rows_per_img >= 32) {
                            float t0[32], t1[32];
                            static_assert(CH == 32 || CH == 16, "chunk");
                            if constexpr (CH == 32) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) { t0[j] = v[j]; t1[j] = v[j] * y[j]; }
                                const float s0 = colsum_group<32>(t0, lane);
                                const float s1 = colsum_group<32>(t1, lane);
                                const int nn = tni * p.nb + (quad * 32) / rows_per_img;
                                if (nn < p.NI) {
                                    atomicAdd(p.stat0 + static_cast<long>(nn) * p.stat_stride + cbase + lane, s0);
                                    atomicAdd(p.stat1 + static_cast<long>(nn) * p.stat_stride + cbase + lane, s1);
                                }
                            }
                        } else {
                            // 16 pixels per image (4x4 maps): half-warp groups.
                            if constexpr (CH == 32) {
#pragma unroll
                                for (int half = 0; half < 2; ++half) {
                                    float t0[16], t1[16];
#pragma unroll
                                    for (int j = 0; j < 16; ++j) { t0[j] = v[half * 16 + j]; t1[j] = v[half * 16 + j] * y[half * 16 + j]; }
                                    const float s0 = colsum_group<16>(t0, lane);
                                    const float s1 = colsum_group<16>(t1, lane);
                                    const int nn = tni * p.nb + (quad * 32 + (lane & 16)) / rows_per_img;
                                    if (nn < p.NI) {
                                        atomicAdd(p.stat0 + static_cast<long>(nn) * p.stat_stride + cbase + half * 16 + (lane & 15), s0);
                                        atomicAdd(p.stat1 + static_cast<long>(nn) * p.stat_stride + cbase + half * 16 + (lane & 15), s1);
                                    }
                                }
                            }
                        }
                    }
                    if (p.aff_a) {
                        const int nn = valid ? n : 0;
                        const float* pa = p.aff_a + static_cast<long>(nn) * p.aff_stride + cbase;
#pragma unroll
                        for (int q = 0; q < CH / 4; ++q) {
                            const float4 a4 = __ldg(reinterpret_cast<const float4*>(pa) + q);
                            v[q * 4 + 0] *= a4.x; v[q * 4 + 1] *= a4.y; v[q * 4 + 2] *= a4.z; v[q * 4 + 3] *= a4.w;
                        }
                    }
                    if (p.addin && valid && cbase < p.addin_climit) {
                        if (!p.addin_pool) {
                            const uint4* src = reinterpret_cast<const uint4*>(p.addin + pix * p.addin_C + cbase);
#pragma unroll
                            for (int q = 0; q < CH / 8; ++q) {
                                const uint4 t = __ldg(src + q);
                                v[q * 8 + 0] += bf16_lo(t.x); v[q * 8 + 1] += bf16_hi(t.x);
                                v[q * 8 + 2] += bf16_lo(t.y); v[q * 8 + 3] += bf16_hi(t.y);
                                v[q * 8 + 4] += bf16_lo(t.z); v[q * 8 + 5] += bf16_hi(t.z);
                                v[q * 8 + 6] += bf16_lo(t.w); v[q * 8 + 7] += bf16_hi(t.w);
                            }
                        } else {
                            const int W2 = p.W * 2, H2 = p.H * 2;
#pragma unroll
                            for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
                                for (int dxx = 0; dxx < 2; ++dxx) {
                                    const long hp = (static_cast<long>(n) * H2 + 2 * h + dy) * W2 + 2 * w + dxx;
                                    const uint4* src = reinterpret_cast<const uint4*>(p.addin + hp * p.addin_C + cbase);
#pragma unroll
                                    for (int q = 0; q < CH / 8; ++q) {
                                        const uint4 t = __ldg(src + q);
                                        v[q * 8 + 0] += bf16_lo(t.x); v[q * 8 + 1] += bf16_hi(t.x);
                                        v[q * 8 + 2] += bf16_lo(t.y); v[q * 8 + 3] += bf16_hi(t.y);
                                        v[q * 8 + 4] += bf16_lo(t.z); v[q * 8 + 5] += bf16_hi(t.z);
                                        v[q * 8 + 6] += bf16_lo(t.w); v[q * 8 + 7] += bf16_hi(t.w);
                                    }
                                }
                            }
                        }
                    }
                    if (valid) {
                        if (p.dx) {
                            uint4* dst = reinterpret_cast<uint4*>(p.dx + pix * p.dx_C + cbase);
#pragma unroll
                            for (int q = 0; q < CH / 8; ++q) {
                                dst[q] = make_uint4(pack_bf16(v[q * 8], v[q * 8 + 1]), pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                                                    pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
                            }
                        }
                        if (p.dx_f32) {
                            float4* dst = reinterpret_cast<float4*>(p.dx_f32 + pix * p.dx_f32_C + cbase);
#pragma unroll
                            for (int q = 0; q < CH / 4; ++q) dst[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                        }
                    }
                }
            }
            // all TMEM reads of this accumulator stage are complete (tmem_ld_wait above)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

}  // namespace p2l
