// StyleGAN2 glue kernels (everything of rosinality's Generator that is not a dense contraction):
// mapping-network GEMVs, style modulation / demodulation, noise + bias + leaky-ReLU, the 4-tap FIR
// blur / up-sampling (upfirdn2d) and the 3-channel ToRGB path, forward and backward.
// The dense 3x3 contractions run on conv_gemm_kernel with SHARED weights: the per-sample weight
// modulation of the reference (ModulatedConv2d: grouped conv with b*Cout filters) is an input-channel
// scale before and an output-channel scale after the convolution.
#include "sg2_kernels.h"

#include <cooperative_groups.h>

#include "conv_gemm.h"

namespace p2l {
namespace cg = cooperative_groups;

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }
__device__ __forceinline__ float b2f(bf16 x) { return a2f(x); }
__device__ __forceinline__ bf16 f2b(float x) { return f2a(x); }
__device__ __forceinline__ void unpack8(const uint4 t, float (&f)[8]) {
    f[0] = act_lo(t.x); f[1] = act_hi(t.x); f[2] = act_lo(t.y); f[3] = act_hi(t.y);
    f[4] = act_lo(t.z); f[5] = act_hi(t.z); f[6] = act_lo(t.w); f[7] = act_hi(t.w);
}
__device__ __forceinline__ uint32_t pk2(float a, float b) { return pack_act(a, b); }
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pk2(f[0], f[1]), pk2(f[2], f[3]), pk2(f[4], f[5]), pk2(f[6], f[7]));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __constant__ float kFir[4] = {0.25f, 0.75f, 0.75f, 0.25f};  // [1,3,3,1]/8 * 2 (per axis)
constexpr float kSqrt2 = 1.41421356237309515f;

// ----------------------------------------------------------------------------- dense latent-side layers
// One small-batch GEMM kernel for every latent-side contraction (mapping network, style affines, demodulation and their
// backward passes):  y[bi][j] = epi( wscale * sum_k pro(x)[bi][k] * M[k * ldm + j] ),  bi < b <= NB.
// Block = 32 outputs j x 8 warps; warp w owns every 8th run of 16 k's, lane = j (coalesced 128-byte rows of M), the
// (transformed) inputs of a 256-wide K chunk sit in shared memory (broadcast reads); the eight partial sums meet in
// shared memory and are added in warp order (no atomics: reproducible).
enum { SG_PRO_ID = 0, SG_PRO_SQUARE = 1, SG_PRO_ACTGRAD = 2, SG_PRO_DEMOD = 3 };
enum { SG_EPI_BIAS_ACT = 0, SG_EPI_STORE = 1, SG_EPI_ACCUM = 2, SG_EPI_SUB_SCALED = 3 };
struct SgGemm {
    const float* x; int ldx;        // inputs [b][K]
    const float* x2; int ldx2;      // PRO_ACTGRAD: y of the forward pass (act' = y > 0 ? 1 : 0.2, * sqrt2); PRO_DEMOD: dm
    const float* M; int ldm;        // [K][ldm]
    const float* bias; float wscale;
    float* y; int ldy;
    const float* s; int lds;        // EPI_SUB_SCALED: y -= s * acc
    int b, K, J, act;
    float* part; int kper;          // split-K (non-batched launches): blockIdx.y owns k in [y * kper, (y+1) * kper) and writes its
                                    // raw sums to part[(y * b + bi) * J + j]; sg_splitk_finish_kernel adds them in y order
};
// `tab` != null: a batch of independent GEMMs in one launch, blockIdx.y picks the descriptor (the per-layer demodulation
// GEMMs and their backward: 15 launches of a few microseconds each become one)
template <int NB>
__global__ void __launch_bounds__(256) sg_gemm_kernel(const SgGemm a0, const SgGemm* __restrict__ tab, int pro, int epi) {
    const SgGemm a = tab ? tab[blockIdx.y] : a0;
    if (blockIdx.x * 32 >= a.J) return;
    constexpr int KC = 128, KW = KC / 8;   // K chunk staged per pass; k's per warp and chunk
    __shared__ float xs[NB][KC];
    __shared__ float red[8][NB][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 32 + lane;
    float acc[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) acc[i] = 0.f;
    const bool split = (tab == nullptr) && a.part != nullptr;
    const int kbeg = split ? blockIdx.y * a.kper : 0, kstop = split ? min(a.K, kbeg + a.kper) : a.K;
    for (int k0 = kbeg; k0 < kstop; k0 += KC) {
        __syncthreads();
        for (int t = threadIdx.x; t < NB * KC; t += 256) {
            const int bi = t / KC, kk = t - bi * KC, k = k0 + kk;
            float v = 0.f;
            if (bi < a.b && k < kstop) {
                v = a.x[(long)bi * a.ldx + k];
                if (pro == SG_PRO_SQUARE) v *= v;
                else if (pro == SG_PRO_ACTGRAD) v *= (a.x2[(long)bi * a.ldx2 + k] > 0.f ? 1.f : 0.2f) * kSqrt2;
                else if (pro == SG_PRO_DEMOD) { const float d = a.x2[(long)bi * a.ldx2 + k]; v *= d * d; }   // x holds ddm * dm
            }
            xs[bi][kk] = v;
        }
        __syncthreads();
        if (j < a.J) {
            const int kend = min(KC, kstop - k0);
            const int kk = warp * KW;   // warp w: k = w*KW .. w*KW+KW-1 of this chunk
            if (kk < kend) {
                const int n = min(KW, kend - kk);
                const float* mp = a.M + (long)(k0 + kk) * a.ldm + j;
#pragma unroll 8
                for (int u = 0; u < n; ++u) {
                    const float w = __ldg(mp + (long)u * a.ldm);
#pragma unroll
                    for (int i = 0; i < NB; ++i) acc[i] = fmaf(xs[i][kk + u], w, acc[i]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) red[warp][i][lane] = acc[i];
    __syncthreads();
    for (int t = threadIdx.x; t < NB * 32; t += 256) {
        const int bi = t >> 5, jj = t & 31, jo = blockIdx.x * 32 + jj;
        if (bi >= a.b || jo >= a.J) continue;
        float v = ((red[0][bi][jj] + red[1][bi][jj]) + (red[2][bi][jj] + red[3][bi][jj])) +
                  ((red[4][bi][jj] + red[5][bi][jj]) + (red[6][bi][jj] + red[7][bi][jj]));
        if (split) {
            a.part[((long)blockIdx.y * a.b + bi) * a.J + jo] = v;
            continue;
        }
        v *= a.wscale;
        float* yp = a.y + (long)bi * a.ldy + jo;
        if (epi == SG_EPI_BIAS_ACT) {
            v += a.bias ? a.bias[jo] : 0.f;
            if (a.act == 1) v = (v > 0.f ? v : 0.2f * v) * kSqrt2;      // fused_leaky_relu
            else if (a.act == 2) v = rsqrtf(v + 1e-8f);                   // demodulation (bias = 0)
            *yp = v;
        } else if (epi == SG_EPI_STORE) {
            *yp = v;
        } else if (epi == SG_EPI_ACCUM) {
            *yp += v;
        } else {
            *yp -= a.s[(long)bi * a.lds + jo] * v;
        }
    }
}
// second stage of a split-K launch: y = epi(wscale * sum_y part[y]) for the store / accumulate epilogues
__global__ void sg_splitk_finish_kernel(const SgGemm a, int ks, int epi) {
    const int jo = blockIdx.x * blockDim.x + threadIdx.x, bi = blockIdx.y;
    if (jo >= a.J) return;
    float v = 0.f;
    for (int y = 0; y < ks; ++y) v += a.part[((long)y * a.b + bi) * a.J + jo];
    v *= a.wscale;
    float* yp = a.y + (long)bi * a.ldy + jo;
    *yp = (epi == SG_EPI_ACCUM) ? *yp + v : v;
}
static void sg_gemm(SgGemm a, int pro, int epi, cudaStream_t st) {
    const int b_all = a.b;
    for (int b0 = 0; b0 < b_all; b0 += 24) {   // <= 24 samples per launch
        SgGemm c = a;
        c.b = b_all - b0 < 24 ? b_all - b0 : 24;
        c.x = a.x + (long)b0 * a.ldx;
        c.x2 = a.x2 ? a.x2 + (long)b0 * a.ldx2 : nullptr;
        c.y = a.y + (long)b0 * a.ldy;
        c.s = a.s ? a.s + (long)b0 * a.lds : nullptr;
        // long K, few output columns (style gradients -> w: K ~ 6000, J = 512): spread K over blocks
        int ks = 1;
        if (c.part && (epi == SG_EPI_STORE || epi == SG_EPI_ACCUM) && a.K >= 2048 && cdiv(a.J, 32) < 64) {
            ks = 8;
            c.kper = ((cdiv(a.K, ks) + 127) / 128) * 128;
            ks = cdiv(a.K, c.kper);
        }
        if (ks <= 1) c.part = nullptr;
        const dim3 grid(cdiv(a.J, 32), ks);
        if (c.b <= 8) sg_gemm_kernel<8><<<grid, 256, 0, st>>>(c, nullptr, pro, epi);
        else if (c.b <= 16) sg_gemm_kernel<16><<<grid, 256, 0, st>>>(c, nullptr, pro, epi);
        else sg_gemm_kernel<24><<<grid, 256, 0, st>>>(c, nullptr, pro, epi);
        count_launch();
        if (ks > 1) {
            sg_splitk_finish_kernel<<<dim3(cdiv(a.J, 128), c.b), 128, 0, st>>>(c, ks, epi);
            count_launch();
        }
    }
}
// ---- batched form: `n` descriptors in device memory (b <= 24 each, same prologue / epilogue mode), one launch
static void sg_gemm_batched(const SgGemm* dtab, int n, int b, int max_J, int pro, int epi, cudaStream_t st) {
    const dim3 grid(cdiv(max_J, 32), n);
    SgGemm none{};
    if (b <= 8) sg_gemm_kernel<8><<<grid, 256, 0, st>>>(none, dtab, pro, epi);
    else if (b <= 16) sg_gemm_kernel<16><<<grid, 256, 0, st>>>(none, dtab, pro, epi);
    else sg_gemm_kernel<24><<<grid, 256, 0, st>>>(none, dtab, pro, epi);
    count_launch();
}
size_t k_sg_batch_bytes(int n) { return (size_t)n * sizeof(SgGemm); }
// descriptor i of a demodulation batch: dm[b, Cout] = rsqrt(sum_i s[b,i]^2 * WsqT[i][o] + 1e-8)
void k_sg_demod_desc(void* host_tab, int i, const float* s, int lds, const float* wsqT, float* dm, int lddm, int b, int Cin, int Cout) {
    SgGemm a{};
    a.x = s; a.ldx = lds; a.M = wsqT; a.ldm = Cout; a.wscale = 1.f; a.y = dm; a.ldy = lddm; a.b = b; a.K = Cin; a.J = Cout; a.act = 2;
    static_cast<SgGemm*>(host_tab)[i] = a;
}
// descriptor i of the demodulation-backward batch (see k_demod_bwd)
void k_sg_demod_bwd_desc(void* host_tab, int i, const float* ddm, const float* dm, int lddm, const float* s, int lds, const float* wsq,
                         float* ds, int ldds, int b, int Cin, int Cout) {
    SgGemm a{};
    a.x = ddm; a.ldx = lddm; a.x2 = dm; a.ldx2 = lddm; a.M = wsq; a.ldm = Cin; a.wscale = 1.f; a.y = ds; a.ldy = ldds;
    a.s = s; a.lds = lds; a.b = b; a.K = Cout; a.J = Cin;
    static_cast<SgGemm*>(host_tab)[i] = a;
}
void k_sg_demod_batched(const void* dev_tab, int n, int b, int max_J, cudaStream_t st) {
    sg_gemm_batched(static_cast<const SgGemm*>(dev_tab), n, b, max_J, SG_PRO_SQUARE, SG_EPI_BIAS_ACT, st);
}
void k_sg_demod_bwd_batched(const void* dev_tab, int n, int b, int max_J, cudaStream_t st) {
    sg_gemm_batched(static_cast<const SgGemm*>(dev_tab), n, b, max_J, SG_PRO_DEMOD, SG_EPI_SUB_SCALED, st);
}
// y[b, j] = act(wscale * sum_k x[b,k] * WT[k][j] + bias[j]); WT row pitch ldw
void k_fc_fwd_ld(const float* x, int ldx, const float* WT, int ldw, const float* bias, float wscale, float* y, int ldy, int b,
                 int in, int out, int act, int square_in, cudaStream_t st) {
    SgGemm a{};
    a.x = x; a.ldx = ldx; a.M = WT; a.ldm = ldw; a.bias = bias; a.wscale = wscale; a.y = y; a.ldy = ldy; a.b = b; a.K = in; a.J = out;
    a.act = act;
    sg_gemm(a, square_in ? SG_PRO_SQUARE : SG_PRO_ID, SG_EPI_BIAS_ACT, st);
}
void k_fc_fwd(const float* x, int ldx, const float* WT, const float* bias, float wscale, float* y, int ldy, int b, int in,
              int out, int act, int square_in, cudaStream_t st) {
    k_fc_fwd_ld(x, ldx, WT, out, bias, wscale, y, ldy, b, in, out, act, square_in, st);
}
// dx[b, k] (+)= wscale * sum_j g[b,j] * W[j][k], g = dy * act'(y); W is [out][in]
void k_fc_bwd(const float* dy, int lddy, const float* y, int ldy, const float* W, float wscale, float* dx, int lddx, int b,
              int in, int out, int act, int accumulate, cudaStream_t st, float* splitk_scratch) {
    SgGemm a{};
    a.part = splitk_scratch;
    a.x = dy; a.ldx = lddy; a.x2 = y; a.ldx2 = ldy; a.M = W; a.ldm = in; a.wscale = wscale; a.y = dx; a.ldy = lddx; a.b = b;
    a.K = out; a.J = in;
    sg_gemm(a, (act == 1 && y) ? SG_PRO_ACTGRAD : SG_PRO_ID, accumulate ? SG_EPI_ACCUM : SG_EPI_STORE, st);
}

// ----------------------------------------------------------------------------- mapping network in ONE launch
// PixelNorm + n_mlp x EqualLinear(512, 512, fused_lrelu) (forward), or the gradient back through them (backward), for
// b <= NB samples: a cluster of 8 CTAs, CTA r owns output columns [64 r, 64 r + 64) of every layer. The layer input
// [NB][512] sits in the shared memory of EVERY CTA; a layer = each warp contracts its 64-wide K slice against the CTA's
// 64 columns of the weight matrix (256-byte coalesced rows, float2 per lane), the eight slices meet in shared memory in
// warp order, and the finished 64 columns are written into the NEXT input buffer of all eight CTAs through distributed
// shared memory; one cluster barrier per layer. 16 launches of ~19 us (latency-bound at 16 blocks) become two.
struct SgMap {
    const float* in;          // forward: z [b][512]; backward: dw [b][512]
    const float* M[8];        // forward: W^T of layer k ([in][out]); backward: W of layer k ([out][in])
    const float* bias[8];     // forward
    float* h[9];              // forward: written (h[0] = pixelnorm(z) .. h[n] = w); backward: read (saved activations)
    const float* z;           // backward: the raw latent (PixelNorm backward)
    float* out;               // backward: dz [b][512]
    const float* row_scale;   // backward: optional per-sample factor
    float wscale, scale;
    int b, n;
};
constexpr int kMapDim = 512, kMapCtas = 8, kMapCols = kMapDim / kMapCtas, kMapWarps = 16, kMapKW = kMapDim / kMapWarps;
template <int NB, bool BWD>
__global__ void __cluster_dims__(kMapCtas, 1, 1) __launch_bounds__(kMapWarps * 32) sg_mapping_kernel(const SgMap a) {
    extern __shared__ __align__(16) float map_sm[];
    float(*xs)[NB][kMapDim] = reinterpret_cast<float(*)[NB][kMapDim]>(map_sm);                       // [2][NB][512]
    float(*red)[NB][kMapCols] = reinterpret_cast<float(*)[NB][kMapCols]>(map_sm + 2 * NB * kMapDim);   // [16][NB][64]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // every layer's 128 KB weight slab of this CTA on its way into L2 (thread = one 256-byte row of each slab)
    for (int k = 0; k < a.n; ++k) {
        const float* row = a.M[k] + (long)threadIdx.x * kMapDim + rank * kMapCols;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 32));
    }
    // ---- the first input, computed by every CTA for itself
    for (int bi = warp; bi < NB; bi += kMapWarps) {
        float v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = bi < a.b ? a.in[(long)bi * kMapDim + q * 32 + lane] : 0.f;
        if constexpr (!BWD) {
            float ss = 0.f;
#pragma unroll
            for (int q = 0; q < 16; ++q) ss += v[q] * v[q];
            ss = warp_sum(ss);
            const float r = rsqrtf(ss / kMapDim + 1e-8f);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                v[q] *= r;
                if (rank == 0 && bi < a.b) a.h[0][(long)bi * kMapDim + q * 32 + lane] = v[q];
            }
        } else {
#pragma unroll
            for (int q = 0; q < 16; ++q)
                if (bi < a.b) v[q] *= (a.h[a.n][(long)bi * kMapDim + q * 32 + lane] > 0.f ? 1.f : 0.2f) * kSqrt2;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) xs[0][bi][q * 32 + lane] = v[q];
    }
    cluster.sync();   // every CTA of the cluster runs (remote shared memory may be written from here on) + local visibility
    for (int L = 0; L < a.n; ++L) {
        const int k = BWD ? a.n - 1 - L : L, cur = L & 1;
        float acc[NB][2];
#pragma unroll
        for (int i = 0; i < NB; ++i) acc[i][0] = acc[i][1] = 0.f;
        const float* mp = a.M[k] + (long)(warp * kMapKW) * kMapDim + rank * kMapCols + lane * 2;
#pragma unroll
        for (int u0 = 0; u0 < kMapKW; u0 += 16) {
            float2 w[16];   // sixteen 256-byte rows in flight per warp
#pragma unroll
            for (int q = 0; q < 16; ++q) w[q] = __ldg(reinterpret_cast<const float2*>(mp + (long)(u0 + q) * kMapDim));
#pragma unroll
            for (int u = 0; u < 16; u += 4) {
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const float4 x4 = *reinterpret_cast<const float4*>(&xs[cur][i][warp * kMapKW + u0 + u]);
                    acc[i][0] = fmaf(x4.x, w[u].x, acc[i][0]); acc[i][1] = fmaf(x4.x, w[u].y, acc[i][1]);
                    acc[i][0] = fmaf(x4.y, w[u + 1].x, acc[i][0]); acc[i][1] = fmaf(x4.y, w[u + 1].y, acc[i][1]);
                    acc[i][0] = fmaf(x4.z, w[u + 2].x, acc[i][0]); acc[i][1] = fmaf(x4.z, w[u + 2].y, acc[i][1]);
                    acc[i][0] = fmaf(x4.w, w[u + 3].x, acc[i][0]); acc[i][1] = fmaf(x4.w, w[u + 3].y, acc[i][1]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) *reinterpret_cast<float2*>(&red[warp][i][lane * 2]) = make_float2(acc[i][0], acc[i][1]);
        __syncthreads();
        for (int t = threadIdx.x; t < NB * kMapCols; t += kMapWarps * 32) {
            const int bi = t / kMapCols, jj = t % kMapCols, jo = rank * kMapCols + jj;
            float s4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                s4[q] = (red[4 * q][bi][jj] + red[4 * q + 1][bi][jj]) + (red[4 * q + 2][bi][jj] + red[4 * q + 3][bi][jj]);
            float v = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * a.wscale;
            if constexpr (!BWD) {
                v += a.bias[k][jo];
                v = (v > 0.f ? v : 0.2f * v) * kSqrt2;
                if (bi < a.b) a.h[k + 1][(long)bi * kMapDim + jo] = v;
            } else {
                if (bi >= a.b) v = 0.f;
                else if (k > 0) v *= (a.h[k][(long)bi * kMapDim + jo] > 0.f ? 1.f : 0.2f) * kSqrt2;
            }
            float* dst = &xs[cur ^ 1][bi][jo];
#pragma unroll
            for (int r = 0; r < kMapCtas; ++r) *cluster.map_shared_rank(dst, r) = v;
        }
        cluster.sync();   // the next input is complete everywhere; nobody still reads the current one or `red`
    }
    if constexpr (BWD) {
        if (rank != 0) return;   // no remote access after the last barrier: leaving is safe
        const float(*g)[kMapDim] = xs[a.n & 1];
        for (int bi = warp; bi < a.b; bi += kMapWarps) {
            float x[16], ss = 0.f, dot = 0.f;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                x[q] = a.z[(long)bi * kMapDim + q * 32 + lane];
                ss += x[q] * x[q];
                dot += x[q] * g[bi][q * 32 + lane];
            }
            ss = warp_sum(ss);
            dot = warp_sum(dot);
            const float r = rsqrtf(ss / kMapDim + 1e-8f);
            const float sc = a.scale * (a.row_scale ? a.row_scale[bi] : 1.f);
#pragma unroll
            for (int q = 0; q < 16; ++q)
                a.out[(long)bi * kMapDim + q * 32 + lane] = sc * (r * g[bi][q * 32 + lane] - x[q] * dot * r * r * r / kMapDim);
        }
    }
}
template <int NB, bool BWD>
static void launch_mapping(const SgMap& a, cudaStream_t st) {
    constexpr size_t smem = (size_t)(2 * NB * kMapDim + kMapWarps * NB * kMapCols) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(sg_mapping_kernel<NB, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    sg_mapping_kernel<NB, BWD><<<kMapCtas, kMapWarps * 32, smem, st>>>(a);
    count_launch();
}
bool k_sg_mapping_fusable(int b, int sdim, int n_mlp) { return sdim == kMapDim && b <= 24 && n_mlp >= 1 && n_mlp <= 8; }
void k_sg_mapping_fwd(const float* z, const float* const* WT, const float* const* bias, float wscale, float* const* h, int b, int n_mlp,
                      cudaStream_t st) {
    SgMap a{};
    a.in = z; a.wscale = wscale; a.b = b; a.n = n_mlp;
    for (int k = 0; k < n_mlp; ++k) { a.M[k] = WT[k]; a.bias[k] = bias[k]; }
    for (int k = 0; k <= n_mlp; ++k) a.h[k] = h[k];
    if (b <= 8) launch_mapping<8, false>(a, st);
    else if (b <= 16) launch_mapping<16, false>(a, st);
    else launch_mapping<24, false>(a, st);
}
void k_sg_mapping_bwd(const float* dw, const float* const* W, float wscale, float* const* h, const float* z, float* dz, float scale,
                      const float* row_scale, int b, int n_mlp, cudaStream_t st) {
    SgMap a{};
    a.in = dw; a.wscale = wscale; a.b = b; a.n = n_mlp; a.z = z; a.out = dz; a.scale = scale; a.row_scale = row_scale;
    for (int k = 0; k < n_mlp; ++k) a.M[k] = W[k];
    for (int k = 0; k <= n_mlp; ++k) a.h[k] = h[k];
    if (b <= 8) launch_mapping<8, true>(a, st);
    else if (b <= 16) launch_mapping<16, true>(a, st);
    else launch_mapping<24, true>(a, st);
}

// PixelNorm: y = x * rsqrt(mean(x^2) + 1e-8); one warp per sample
__global__ void pixelnorm_fwd_kernel(const float* x, float* y, int n) {
    const int bi = blockIdx.x, lane = threadIdx.x;
    float ss = 0.f;
    for (int k = lane; k < n; k += 32) ss += x[bi * n + k] * x[bi * n + k];
    ss = warp_sum(ss);
    const float r = rsqrtf(ss / n + 1e-8f);
    for (int k = lane; k < n; k += 32) y[bi * n + k] = x[bi * n + k] * r;
}
__global__ void pixelnorm_bwd_kernel(const float* x, const float* dy, float* dx, int n, float scale, const float* row_scale) {
    const int bi = blockIdx.x, lane = threadIdx.x;
    float ss = 0.f, dot = 0.f;
    for (int k = lane; k < n; k += 32) {
        ss += x[bi * n + k] * x[bi * n + k];
        dot += x[bi * n + k] * dy[bi * n + k];
    }
    ss = warp_sum(ss);
    dot = warp_sum(dot);
    const float r = rsqrtf(ss / n + 1e-8f);
    const float sc = scale * (row_scale ? row_scale[bi] : 1.f);
    // d/dx_k [x_j r] = r delta_jk - x_j x_k r^3 / n
    for (int k = lane; k < n; k += 32) dx[bi * n + k] = sc * (r * dy[bi * n + k] - x[bi * n + k] * dot * r * r * r / n);
}
void k_pixelnorm_fwd(const float* x, float* y, int b, int n, cudaStream_t st) {
    pixelnorm_fwd_kernel<<<b, 32, 0, st>>>(x, y, n); count_launch();
}
void k_pixelnorm_bwd(const float* x, const float* dy, float* dx, int b, int n, float scale, const float* row_scale, cudaStream_t st) {
    pixelnorm_bwd_kernel<<<b, 32, 0, st>>>(x, dy, dx, n, scale, row_scale); count_launch();
}

// ds[b,i] += -s[b,i] * sum_o (ddm*dm^3)[b,o] * Wsq[o][i]   (gradient through the demodulation dm = rsqrt(sum s^2 Wsq + eps));
// `ddm` arrives MULTIPLIED by dm (the producers sum g * u * dm per column and leave the division to this kernel)
void k_demod_bwd(const float* ddm, const float* dm, int lddm, const float* s, int lds, const float* Wsq, float* ds, int ldds,
                 int b, int Cin, int Cout, cudaStream_t st) {
    SgGemm a{};
    a.x = ddm; a.ldx = lddm; a.x2 = dm; a.ldx2 = lddm; a.M = Wsq; a.ldm = Cin; a.wscale = 1.f; a.y = ds; a.ldy = ldds;
    a.s = s; a.lds = lds; a.b = b; a.K = Cout; a.J = Cin;
    sg_gemm(a, SG_PRO_DEMOD, SG_EPI_SUB_SCALED, st);
}

// ----------------------------------------------------------------------------- modulation
// A[b, p, c] = x[b or 0, p, c] * s[b, c] (layer 0: the learned constant times its style; every later layer's
// modulated input comes out of the previous layer's convolution epilogue)
__global__ void modulate_kernel(const bf16* __restrict__ x, long x_bstride, const float* __restrict__ s, int lds, bf16* __restrict__ A,
                                int b, int H, int W, int C) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int CG = C / 8;
    const long total = (long)b * H * W * CG;
    if (i >= total) return;
    const int cg = i % CG;
    const long q = i / CG;
    const long pp = q % ((long)W * H);
    const int bi = q / ((long)W * H);
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + bi * x_bstride + pp * C + cg * 8)), v);
    const float4 s0 = *reinterpret_cast<const float4*>(s + (long)bi * lds + cg * 8);
    const float4 s1 = *reinterpret_cast<const float4*>(s + (long)bi * lds + cg * 8 + 4);
    v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w; v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
    *reinterpret_cast<uint4*>(A + q * C + cg * 8) = pack8(v);
}
void k_sg_modulate(const bf16* x, long x_bstride, const float* s, int lds, bf16* A, int b, int H, int W, int C, cudaStream_t st) {
    const long total = (long)b * H * W * (C / 8);
    modulate_kernel<<<cdiv(total, 256), 256, 0, st>>>(x, x_bstride, s, lds, A, b, H, W, C); count_launch();
}

// backward: ds[b,c] += sum_p dA*x ; dx = s*dA (when dx != null)
// block = 8 channel groups x 32 pixels, one image; 64 channels per blockIdx.y
__global__ void modulate_bwd_kernel(const bf16* __restrict__ dA, const bf16* __restrict__ x, long x_bstride, const float* __restrict__ s,
                                    int lds, bf16* __restrict__ dx, float* __restrict__ part, int H, int W, int C) {
    __shared__ float red[32][65];
    const int cg = threadIdx.x, py = threadIdx.y;
    const int c = blockIdx.y * 64 + cg * 8;
    const int bi = blockIdx.z;
    const int HW = H * W;
    float sv[8], acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sv[e] = s[(long)bi * lds + c + e]; acc[e] = 0.f; }
    for (int p = blockIdx.x * 256 + py; p < min(HW, (int)(blockIdx.x + 1) * 256); p += 32) {
        const long ia = ((long)bi * HW + p) * C + c;
        float g[8], xv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dA + ia)), g);
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + bi * x_bstride + (long)p * C + c)), xv);
#pragma unroll
        for (int e = 0; e < 8; ++e) { acc[e] += g[e] * xv[e]; g[e] *= sv[e]; }
        if (dx) *reinterpret_cast<uint4*>(dx + ((long)bi * HW + p) * C + c) = pack8(g);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[py][cg * 8 + e] = acc[e];
    __syncthreads();
    const int tid = py * 8 + cg;
    if (tid < 64) {
        float t = 0.f;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) t += red[k][tid];
        part[((long)bi * gridDim.x + blockIdx.x) * C + blockIdx.y * 64 + tid] = t;  // one writer per slot (no atomics)
    }
}
// dst[bi][j] += sum over the `parts` pixel blocks of part[bi][p][j], in block order: the reproducible second stage of the
// per-(sample, channel) reductions below
__global__ void __launch_bounds__(512) partial_reduce_add_kernel(const float* __restrict__ part, int parts, int n,
                                                                 float* __restrict__ dst, int ld) {
    __shared__ float r[16][33];
    const int j = blockIdx.x * 32 + threadIdx.x, bi = blockIdx.y, sl = threadIdx.y;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    if (j < n) {
        const float* base = part + (long)bi * parts * n + j;
        int q = sl;
        for (; q + 48 < parts; q += 64) {
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] += __ldg(base + (long)(q + 16 * u) * n);
        }
#pragma unroll
        for (int u = 0; u < 3; ++u)
            if (q + 16 * u < parts) a[u] += __ldg(base + (long)(q + 16 * u) * n);
    }
    r[sl][threadIdx.x] = (a[0] + a[1]) + (a[2] + a[3]);
    __syncthreads();
    for (int h = 8; h > 0; h >>= 1) {
        if (sl < h) r[sl][threadIdx.x] += r[sl + h][threadIdx.x];
        __syncthreads();
    }
    if (sl == 0 && j < n) dst[(long)bi * ld + j] += r[0][threadIdx.x];
}
static void partial_reduce_add(const float* part, int parts, int b, int n, float* dst, int ld, cudaStream_t st) {
    partial_reduce_add_kernel<<<dim3(cdiv(n, 32), b), dim3(32, 16), 0, st>>>(part, parts, n, dst, ld); count_launch();
}
long k_sg_scratch_floats(int b, int H, int W, int C) { return (long)b * cdiv((long)H * W, 256) * C * 4; }
void k_sg_modulate_bwd(const bf16* dA, const bf16* x, long x_bstride, const float* s, int lds, bf16* dx, float* ds, int ldds,
                       float* scratch, int b, int H, int W, int C, cudaStream_t st) {
    dim3 grid(cdiv((long)H * W, 256), C / 64, b), block(8, 32);
    modulate_bwd_kernel<<<grid, block, 0, st>>>(dA, x, x_bstride, s, lds, dx, scratch, H, W, C); count_launch();
    partial_reduce_add(scratch, (int)grid.x, b, C, ds, ldds, st);
}

// ----------------------------------------------------------------------------- last layer's activation backward
// The gradient through layer l's leaky-ReLU / noise / bias / demodulation normally lives in the epilogue of layer l+1's
// dgrad (sg_epilogue.cuh); the LAST layer has no successor, so it runs here as an element-wise pass with the same
// arithmetic: g = dx * sqrt2 * lrelu'(x) ; u * dm = lrelu^-1(x / sqrt2) - nw * noise - bias ; (ddm * dm)[b,c] += sum_p g * u * dm ;
// G = dm * g (16-bit). block = 8 channel groups x 32 pixels, 64 channels per blockIdx.y
__global__ void post_bwd_x_kernel(const bf16* __restrict__ dx, const bf16* __restrict__ x, const float* __restrict__ dm, int lddm,
                                  const float* __restrict__ noise, const float* __restrict__ nw, const float* __restrict__ bias,
                                  bf16* __restrict__ G, float* __restrict__ part, int H, int W, int C) {
    __shared__ float red[32][65];
    const int cg = threadIdx.x, py = threadIdx.y;
    const int c = blockIdx.y * 64 + cg * 8;
    const int bi = blockIdx.z;
    const int HW = H * W;
    float dmv[8], bv[8], acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { dmv[e] = dm[(long)bi * lddm + c + e]; bv[e] = bias[c + e]; acc[e] = 0.f; }
    const float nwv = noise ? nw[0] : 0.f;
    for (int p = blockIdx.x * 256 + py; p < min(HW, (int)(blockIdx.x + 1) * 256); p += 32) {
        const long o = ((long)bi * HW + p) * C + c;
        float g[8], xv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dx + o)), g);
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + o)), xv);
        const float nz = noise ? nwv * noise[(long)bi * HW + p] : 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const bool pos = xv[e] > 0.f;
            g[e] *= pos ? kSqrt2 : 0.2f * kSqrt2;
            const float pre = xv[e] * (pos ? (1.f / kSqrt2) : (1.f / (0.2f * kSqrt2)));
            acc[e] += g[e] * (pre - nz - bv[e]);   // ddm * dm (see k_demod_bwd)
            g[e] *= dmv[e];
        }
        *reinterpret_cast<uint4*>(G + o) = pack8(g);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[py][cg * 8 + e] = acc[e];
    __syncthreads();
    const int tid = py * 8 + cg;
    if (tid < 64) {
        float t = 0.f;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) t += red[k][tid];
        part[((long)bi * gridDim.x + blockIdx.x) * C + blockIdx.y * 64 + tid] = t;
    }
}
void k_sg_post_bwd_x(const bf16* dx, const bf16* x, const float* dm, int lddm, const float* noise, const float* nw, const float* bias,
                     bf16* G, float* ddm, float* scratch, int b, int H, int W, int C, cudaStream_t st) {
    dim3 grid(cdiv((long)H * W, 256), C / 64, b), block(8, 32);
    post_bwd_x_kernel<<<grid, block, 0, st>>>(dx, x, dm, lddm, noise, nw, bias, G, scratch, H, W, C); count_launch();
    partial_reduce_add(scratch, (int)grid.x, b, C, ddm, lddm, st);
}

// ----------------------------------------------------------------------------- ToRGB
// weff[b, c, i] = Wr[c, i] * s[b, i] * scale
__global__ void weff_kernel(const float* Wr, const float* s, int lds, float scale, float* weff, int b, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b * 3 * C) return;
    const int k = i % C, c = (i / C) % 3, bi = i / (3 * C);
    weff[i] = Wr[c * C + k] * s[(long)bi * lds + k] * scale;
}
void k_sg_weff(const float* Wr, const float* s, int lds, float scale, float* weff, int b, int C, cudaStream_t st) {
    weff_kernel<<<cdiv((long)b * 3 * C, 256), 256, 0, st>>>(Wr, s, lds, scale, weff, b, C); count_launch();
}
// FIR x2 up-sampling of the previous rgb (upfirdn2d up=2, pad (2,1)): out[y] = sum_{t: (y+t) even} prev[(y+t-2)/2] k[t].
// Per axis that is two taps: rows ia = floor((y-1)/2) and ia+1 with weights (1/4, 3/4) for even y, (3/4, 1/4) for odd y.
__device__ __forceinline__ float up_gather(const float* __restrict__ prev, int h, int w, int y, int xx) {
    const int ia = (y - 1) >> 1, ja = (xx - 1) >> 1;
    const float ya = (y & 1) ? 0.75f : 0.25f, xa = (xx & 1) ? 0.75f : 0.25f;
    // out-of-range taps: clamped address, zero weight — four independent loads in flight, same sum order as the tap loop
    const float wya = ia >= 0 ? ya : 0.f, wyb = ia + 1 < h ? 1.f - ya : 0.f;
    const float wxa = ja >= 0 ? xa : 0.f, wxb = ja + 1 < w ? 1.f - xa : 0.f;
    const float* r0 = prev + (long)max(ia, 0) * w;
    const float* r1 = prev + (long)min(ia + 1, h - 1) * w;
    const int j0 = max(ja, 0), j1 = min(ja + 1, w - 1);
    const float v00 = __ldg(r0 + j0), v01 = __ldg(r0 + j1), v10 = __ldg(r1 + j0), v11 = __ldg(r1 + j1);
    float acc = 0.f;
    acc = fmaf(wya * wxa, v00, acc);
    acc = fmaf(wya * wxb, v01, acc);
    acc = fmaf(wyb * wxa, v10, acc);
    acc = fmaf(wyb * wxb, v11, acc);
    return acc;
}
// rgb[b,c,p] = sum_i weff[b,c,i] x[b,p,i] + bias[c] (+ upsampled previous rgb). block = 256 consecutive pixels of one image,
// 32 per warp. A warp load covers 512 contiguous bytes of x (32/lpp pixels, lpp = min(32, C/8) lanes per pixel, 8 channels per
// lane; C = 512 takes two loads per pixel), the lane's 3 x 8 weights sit in registers, the three sums are reduced over the lanes
// of a pixel by shuffles and leave through shared memory so that the planar stores are coalesced.
template <int NK>
__global__ void __launch_bounds__(256) torgb_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ weff,
                                                        const float* __restrict__ bias, const float* __restrict__ prev,
                                                        float* __restrict__ rgb, int H, int W, int C, int tiles_per_block) {
    __shared__ float res[3][256];
    const int bi = blockIdx.y, HW = H * W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lpp = min(32, C >> 3), ppl = 32 / lpp;
    const int sub = lane % lpp, cl = sub * 8;
    float w[NK][3][8];
#pragma unroll
    for (int k = 0; k < NK; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4* wp = reinterpret_cast<const float4*>(weff + ((long)bi * 3 + c) * C + k * 256 + cl);
            const float4 lo = __ldg(wp), hi = __ldg(wp + 1);
            w[k][c][0] = lo.x; w[k][c][1] = lo.y; w[k][c][2] = lo.z; w[k][c][3] = lo.w;
            w[k][c][4] = hi.x; w[k][c][5] = hi.y; w[k][c][6] = hi.z; w[k][c][7] = hi.w;
        }
    const float b0 = __ldg(bias), b1 = __ldg(bias + 1), b2 = __ldg(bias + 2);
    const bf16* xb = x + (long)bi * HW * C + cl;
    for (int tile = 0; tile < tiles_per_block; ++tile) {
        const int p0 = (blockIdx.x * tiles_per_block + tile) * 256;
        if (p0 >= HW) break;
        const int pw = p0 + warp * 32 + lane / lpp;
        for (int q0 = 0; q0 < 32; q0 += 8 * ppl) {
            uint4 raw[8][NK];   // eight warp loads (x NK) in flight before the first use
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int p = pw + q0 + j * ppl;
#pragma unroll
                for (int k = 0; k < NK; ++k)
                    raw[j][k] = p < HW ? __ldg(reinterpret_cast<const uint4*>(xb + (long)p * C + k * 256)) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int p = pw + q0 + j * ppl;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                    float v[8];
                    unpack8(raw[j][k], v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        a0 = fmaf(v[e], w[k][0][e], a0);
                        a1 = fmaf(v[e], w[k][1][e], a1);
                        a2 = fmaf(v[e], w[k][2][e], a2);
                    }
                }
#pragma unroll
                for (int h = 16; h > 0; h >>= 1) {
                    if (h < lpp) {
                        a0 += __shfl_xor_sync(0xffffffffu, a0, h);
                        a1 += __shfl_xor_sync(0xffffffffu, a1, h);
                        a2 += __shfl_xor_sync(0xffffffffu, a2, h);
                    }
                }
                if (sub == 0 && p < HW) { res[0][p - p0] = a0; res[1][p - p0] = a1; res[2][p - p0] = a2; }
            }
        }
        __syncthreads();
        const int p = p0 + threadIdx.x;   // thread = pixel: one division for the three planes
        if (p < HW) {
            float v0 = res[0][threadIdx.x] + b0, v1 = res[1][threadIdx.x] + b1, v2 = res[2][threadIdx.x] + b2;
            if (prev) {
                const int y = p / W, xx = p - y * W, h2 = H >> 1, w2 = W >> 1;
                const float* pp = prev + (long)bi * 3 * h2 * w2;
                v0 += up_gather(pp, h2, w2, y, xx);
                v1 += up_gather(pp + h2 * w2, h2, w2, y, xx);
                v2 += up_gather(pp + 2 * h2 * w2, h2, w2, y, xx);
            }
            float* o = rgb + (long)bi * 3 * HW + p;
            o[0] = v0; o[HW] = v1; o[2L * HW] = v2;
        }
        __syncthreads();
    }
}
// any other channel count (multiple of 8): thread per pixel, the three weight rows in shared memory
__global__ void __launch_bounds__(128) torgb_fwd_generic_kernel(const bf16* __restrict__ x, const float* __restrict__ weff,
                                                        const float* __restrict__ bias, const float* __restrict__ prev,
                                                        float* __restrict__ rgb, int H, int W, int C) {
    extern __shared__ float sw[];  // [3][C]
    const int bi = blockIdx.y;
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) sw[i] = weff[(long)bi * 3 * C + i];
    __syncthreads();
    const int HW = H * W;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
        const uint4* xp = reinterpret_cast<const uint4*>(x + ((long)bi * HW + p) * C);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int k = 0; k < C; k += 8) {
            float v[8];
            unpack8(__ldg(xp + (k >> 3)), v);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                a0 = fmaf(v[e], sw[k + e], a0);
                a1 = fmaf(v[e], sw[C + k + e], a1);
                a2 = fmaf(v[e], sw[2 * C + k + e], a2);
            }
        }
        const int y = p / W, xx = p % W;
        const float acc[3] = {a0, a1, a2};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = acc[c] + bias[c];
            if (prev) v += up_gather(prev + ((long)bi * 3 + c) * (H / 2) * (W / 2), H / 2, W / 2, y, xx);
            rgb[((long)bi * 3 + c) * HW + p] = v;
        }
    }
}
void k_sg_torgb_fwd(const bf16* x, const float* weff, const float* bias, const float* prev, float* rgb, int b, int H, int W, int C,
                    cudaStream_t st) {
    // several 256-pixel tiles per block at the large resolutions: the per-lane weights are fetched once per block
    const int tiles = cdiv((long)H * W, 256), tpb = tiles >= 1024 ? 4 : (tiles >= 256 ? 2 : 1);
    const dim3 grid(cdiv(tiles, tpb), b);
    if (C < 64 || C > 512 || (C & (C - 1))) {
        int gx = cdiv((long)H * W, 128);
        if (gx > 1024) gx = 1024;
        torgb_fwd_generic_kernel<<<dim3(gx, b), 128, (size_t)3 * C * sizeof(float), st>>>(x, weff, bias, prev, rgb, H, W, C);
    } else if (C <= 256) torgb_fwd_kernel<1><<<grid, 256, 0, st>>>(x, weff, bias, prev, rgb, H, W, C, tpb);
    else torgb_fwd_kernel<2><<<grid, 256, 0, st>>>(x, weff, bias, prev, rgb, H, W, C, tpb);
    count_launch();
}
// backward: dx[b,p,i] (+)= sum_c drgb[b,c,p] weff[b,c,i] ; dweff[b,c,i] += sum_p drgb[b,c,p] x[b,p,i]
// block = 8 channel groups x 32 pixels, 256 pixels per block (eight independent iterations per thread: loads in flight).
// POST (the LAST layer, whose output feeds only its ToRGB): the activation / noise / bias / demodulation backward of
// post_bwd_x_kernel runs on the gradient while it is still in registers — G = dm * g is stored instead of dx, and the
// (ddm * dm) sums leave through a fourth plane of partial slots.
template <bool POST>
__global__ void __launch_bounds__(256, 2) torgb_bwd_kernel(const float* __restrict__ drgb, const bf16* __restrict__ x,
                                                           const float* __restrict__ weff, bf16* __restrict__ dx, float* __restrict__ part,
                                                           int H, int W, int C, int accumulate, const float* __restrict__ dm, int lddm,
                                                           const float* __restrict__ noise, const float* __restrict__ nw,
                                                           const float* __restrict__ bias, float* __restrict__ part_dm) {
    __shared__ float red[POST ? 4 : 3][32][65];
    __shared__ __align__(16) float coef[5][64];   // the block's 64 channels: three weight rows, dm, bias (registers go to the sums)
    const int cg = threadIdx.x, py = threadIdx.y;
    const int c = blockIdx.y * 64 + cg * 8;
    const int bi = blockIdx.z;
    const int HW = H * W;
    const int tid = py * 8 + cg;
    if (tid < 192) coef[tid >> 6][tid & 63] = weff[((long)bi * 3 + (tid >> 6)) * C + blockIdx.y * 64 + (tid & 63)];
    else if (POST) {
        coef[3][tid & 63] = dm[(long)bi * lddm + blockIdx.y * 64 + (tid & 63)];
        coef[4][tid & 63] = bias[blockIdx.y * 64 + (tid & 63)];
    }
    __syncthreads();
    float a0[8], a1[8], a2[8], a3[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a0[e] = a1[e] = a2[e] = a3[e] = 0.f;
    const float nwv = (POST && noise) ? nw[0] : 0.f;
    auto row8 = [&](int r, float (&v)[8]) {
        const float4 lo = *reinterpret_cast<const float4*>(&coef[r][cg * 8]), hi = *reinterpret_cast<const float4*>(&coef[r][cg * 8 + 4]);
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    };
    constexpr int B = 4;   // iterations whose loads are issued together
    for (int it0 = 0; it0 < 8; it0 += B) {
        uint4 xr[B], dr[B];
        float g[B][3], nzv[B];
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = blockIdx.x * 256 + (it0 + j) * 32 + py;
            const bool ok = p < HW;
            const long o = ((long)bi * HW + p) * C + c;
            xr[j] = ok ? __ldg(reinterpret_cast<const uint4*>(x + o)) : make_uint4(0u, 0u, 0u, 0u);
            dr[j] = (!POST && accumulate && ok) ? *reinterpret_cast<const uint4*>(dx + o) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) g[j][cc] = ok ? __ldg(drgb + ((long)bi * 3 + cc) * HW + p) : 0.f;
            nzv[j] = (POST && noise && ok) ? nwv * __ldg(noise + (long)bi * HW + p) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int p = blockIdx.x * 256 + (it0 + j) * 32 + py;
            const long o = ((long)bi * HW + p) * C + c;
            float xv[8], d[8], w[8];
            unpack8(xr[j], xv);
            unpack8(dr[j], d);
            const float g0 = g[j][0], g1 = g[j][1], g2 = g[j][2], nz = nzv[j];
            row8(0, w);
#pragma unroll
            for (int e = 0; e < 8; ++e) { d[e] = fmaf(g0, w[e], d[e]); a0[e] = fmaf(g0, xv[e], a0[e]); }
            row8(1, w);
#pragma unroll
            for (int e = 0; e < 8; ++e) { d[e] = fmaf(g1, w[e], d[e]); a1[e] = fmaf(g1, xv[e], a1[e]); }
            row8(2, w);
#pragma unroll
            for (int e = 0; e < 8; ++e) { d[e] = fmaf(g2, w[e], d[e]); a2[e] = fmaf(g2, xv[e], a2[e]); }
            if constexpr (POST) {
                float bv[8];
                row8(4, bv);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const bool pos = xv[e] > 0.f;
                    d[e] *= pos ? kSqrt2 : 0.2f * kSqrt2;
                    const float pre = xv[e] * (pos ? (1.f / kSqrt2) : (1.f / (0.2f * kSqrt2)));
                    a3[e] += d[e] * (pre - nz - bv[e]);   // ddm * dm (see k_demod_bwd)
                }
                row8(3, w);
#pragma unroll
                for (int e = 0; e < 8; ++e) d[e] *= w[e];
            }
            if (p < HW) *reinterpret_cast<uint4*>(dx + o) = pack8(d);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        red[0][py][cg * 8 + e] = a0[e]; red[1][py][cg * 8 + e] = a1[e]; red[2][py][cg * 8 + e] = a2[e];
        if constexpr (POST) red[3][py][cg * 8 + e] = a3[e];
    }
    __syncthreads();
    if (tid < (POST ? 256 : 192)) {
        const int cc = tid / 64, k = tid % 64;
        float t = 0.f;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) t += red[cc][j][k];
        if (cc < 3) part[(((long)bi * gridDim.x + blockIdx.x) * 3 + cc) * C + blockIdx.y * 64 + k] = t;
        else part_dm[((long)bi * gridDim.x + blockIdx.x) * C + blockIdx.y * 64 + k] = t;
    }
}
void k_sg_torgb_bwd(const float* drgb, const bf16* x, const float* weff, bf16* dx, float* dweff, float* scratch, int b, int H, int W,
                    int C, int accumulate, cudaStream_t st) {
    dim3 grid(cdiv((long)H * W, 256), C / 64, b), block(8, 32);
    torgb_bwd_kernel<false><<<grid, block, 0, st>>>(drgb, x, weff, dx, scratch, H, W, C, accumulate, nullptr, 0, nullptr, nullptr,
                                                    nullptr, nullptr);
    count_launch();
    partial_reduce_add(scratch, (int)grid.x, b, 3 * C, dweff, 3 * C, st);
}
// the last layer: ToRGB backward + that layer's activation backward in one pass over x (see torgb_bwd_kernel<true>)
void k_sg_torgb_post_bwd(const float* drgb, const bf16* x, const float* weff, float* dweff, const float* dm, int lddm, const float* noise,
                         const float* nw, const float* bias, bf16* G, float* ddm, float* scratch, int b, int H, int W, int C,
                         cudaStream_t st) {
    dim3 grid(cdiv((long)H * W, 256), C / 64, b), block(8, 32);
    float* scratch_dm = scratch + (long)b * grid.x * 3 * C;
    torgb_bwd_kernel<true><<<grid, block, 0, st>>>(drgb, x, weff, G, scratch, H, W, C, 0, dm, lddm, noise, nw, bias, scratch_dm);
    count_launch();
    partial_reduce_add(scratch, (int)grid.x, b, 3 * C, dweff, 3 * C, st);
    partial_reduce_add(scratch_dm, (int)grid.x, b, C, ddm, lddm, st);
}
// ds[b, i] += scale * sum_c dweff[b,c,i] * Wr[c,i]
__global__ void weff_bwd_kernel(const float* dweff, const float* Wr, float scale, float* ds, int ldds, int b, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b * C) return;
    const int k = i % C, bi = i / C;
    float t = 0.f;
    for (int c = 0; c < 3; ++c) t += dweff[((long)bi * 3 + c) * C + k] * Wr[c * C + k];
    ds[(long)bi * ldds + k] += t * scale;
}
void k_sg_weff_bwd(const float* dweff, const float* Wr, float scale, float* ds, int ldds, int b, int C, cudaStream_t st) {
    weff_bwd_kernel<<<cdiv((long)b * C, 256), 256, 0, st>>>(dweff, Wr, scale, ds, ldds, b, C); count_launch();
}
// adjoint of the rgb up-sampling: dprev[i, j] = sum_{t,v} drgb[2i + 2 - t, 2j + 2 - v] k[t] k[v]
__global__ void rgb_up_adjoint_kernel(const float* __restrict__ drgb, float* __restrict__ dprev, int b3, int h, int w) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)b3 * h * w) return;
    const int j = idx % w, i = (idx / w) % h;
    const long pl = idx / ((long)w * h);
    const int H = 2 * h, W = 2 * w;
    const float* src = drgb + pl * H * W;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int y = 2 * i + 2 - t;
        if (y < 0 || y >= H) continue;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int xx = 2 * j + 2 - v;
            if (xx < 0 || xx >= W) continue;
            acc = fmaf(kFir[t] * kFir[v], src[(long)y * W + xx], acc);
        }
    }
    dprev[idx] = acc;
}
void k_sg_rgb_up_adjoint(const float* drgb, float* dprev, int b, int h, int w, cudaStream_t st) {
    rgb_up_adjoint_kernel<<<cdiv((long)b * 3 * h * w, 256), 256, 0, st>>>(drgb, dprev, b * 3, h, w); count_launch();
}
// img = clamp(rgb, -1, 1)
__global__ void clamp_kernel(const float* rgb, float* img, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) img[i] = fminf(fmaxf(rgb[i], -1.f), 1.f);
}
void k_sg_clamp(const float* rgb, float* img, long n, cudaStream_t st) { clamp_kernel<<<cdiv(n, 256), 256, 0, st>>>(rgb, img, n); count_launch(); }
// drgb = dimg where -1 <= rgb <= 1 else 0
__global__ void clamp_bwd_kernel(const float* rgb, const float* dimg, float* drgb, long n, float scale) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) drgb[i] = (rgb[i] >= -1.f && rgb[i] <= 1.f) ? dimg[i] * scale : 0.f;
}
void k_sg_clamp_bwd(const float* rgb, const float* dimg, float* drgb, long n, float scale, cudaStream_t st) {
    clamp_bwd_kernel<<<cdiv(n, 256), 256, 0, st>>>(rgb, dimg, drgb, n, scale); count_launch();
}

// ----------------------------------------------------------------------------- w+ / noise search (stylegan2.py:122-125)
// NoiseInjection backward: x = lrelu(dm*u + nw*noise + bias)*sqrt2  =>  dnoise[b,p] = nw * sum_c dx[b,p,c]*sqrt2*lrelu'(x[b,p,c]);
// one warp per pixel; `scale` (and row_scale[b]) remove the 16-bit gradient scale / apply the upstream factor
__global__ void noise_bwd_kernel(const bf16* __restrict__ dx, const bf16* __restrict__ x, const float* __restrict__ nw,
                                 float* __restrict__ dnoise, long npix, int HW, int C, float scale, const float* __restrict__ row_scale) {
    const long p = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= npix) return;
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
        float g[8], xv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dx + p * C + c)), g);
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + p * C + c)), xv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc += g[e] * (xv[e] > 0.f ? 1.f : 0.2f);
    }
    acc = warp_sum(acc);
    if (lane == 0) dnoise[p] = acc * kSqrt2 * nw[0] * scale * (row_scale ? row_scale[p / HW] : 1.f);
}
void k_sg_noise_bwd(const bf16* dx, const bf16* x, const float* nw, float* dnoise, int b, int H, int W, int C, float scale,
                    const float* row_scale, cudaStream_t st) {
    const long npix = (long)b * H * W;
    noise_bwd_kernel<<<cdiv(npix, 8), 256, 0, st>>>(dx, x, nw, dnoise, npix, H * W, C, scale, row_scale); count_launch();
}
// x[b, n] *= scale * (row_scale ? row_scale[b] : 1)
__global__ void scale_out_kernel(float* x, long n, float scale, const float* row_scale) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[(long)blockIdx.y * n + i] *= scale * (row_scale ? row_scale[blockIdx.y] : 1.f);
}
void k_sg_scale_out(float* x, int b, long n, float scale, const float* row_scale, cudaStream_t st) {
    dim3 grid(cdiv(n, 256), b);
    scale_out_kernel<<<grid, 256, 0, st>>>(x, n, scale, row_scale); count_launch();
}

}  // namespace p2l
