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
constexpr int kStatRedBytes = 4096;  // cross-warp staging of the BN-gradient sums: 2 groups x 2 parities x [4][2][32] floats
#ifndef P2L_OCC
#define P2L_OCC 2
#endif
enum { EPI_FWD = 0, EPI_BWD = 1 };
// epilogue flavours: every flavour is its own kernel instantiation, so that no kernel carries another one's epilogue code
// (the epilogue body is the hot loop of 8 warps per SM; its size is what the instruction cache sees)
enum { FLAVOR_PLAIN = 0, FLAVOR_ROWFUSE = 1, FLAVOR_SG = 2 };

struct ConvGemmParams {
    // ---- M: output pixel grid [NI, H, W], tile box (nb, th, tw), tw*th*nb == 128
    int NI, H, W;
    int tw, th, nb;   // powers of two
    int ltw, lth;     // log2(tw), log2(th)
    int tiles_w, tiles_h, tiles_n;
    // ---- N
    int Cout;     // logical output channels
    int n_tiles;  // ceil(Cout / BN)
    // ---- K
    int taps_h, taps_w, pad_h, pad_w;
    int cin_chunks;  // Cin / 64
    int a_c0;        // channel offset into the A tensor
    int b_batched;   // B tensor map's 3rd coordinate = image index
    int halo_bo;     // halo kernel: fill the descriptor's base-offset field from the start address
    int halo_resb;   // halo kernel: the whole weight matrix (9 * cin_chunks tiles) stays resident in shared memory
    int halo_sa;     // halo kernel: A patch ring depth
    int halo_sb;     // halo kernel: B tile ring depth (non-resident)
    // ---- epilogue, forward
    float alpha;             // scale on the accumulator
    const float* alpha_ptr;  // optional device scalar multiplied into alpha (attention gamma); both modes
    const float* bias;       // [Cout] or null
    const act_t* resid;  // skip input, NHWC [NI, H>>resid_shift, W>>resid_shift, resid_C]
    int resid_C, resid_shift;
    act_t* raw;  // raw (pre-affine) output, NHWC, channel stride raw_C
    int raw_C;
    float* raw_f32;  // same, fp32 (attention logits)
    int raw_f32_C;
    const float* aff_a;  // per-sample affine [NI, aff_stride] (pointer pre-offset to this layer)
    const float* aff_s;
    int aff_stride;
    int relu;
    act_t* act;  // activated output, NHWC, channel stride act_C
    int act_C;
    int act_up;             // write act to the 2x nearest-upsampled grid [NI, 2H, 2W, act_C]
    act_t* act_lo;  // with act_up: also keep the low-res copy (needed by backward)
    float* img_nchw;        // tanh(v) for c < Cout written as fp32 NCHW [NI, Cout, H, W]
    int img_linear;         // ... without the tanh (planar fp32 output of a tap-expanded head)
    // ---- epilogue, forward: row-wise softmax / softmax-gradient fusions (attention, "attn_fused" option; ROWFUSE
    //      instantiations of the direct-epilogue kernel only). A thread owns one accumulator row, so row reductions are thread-local.
    float* rowstat;          // pass 1 of a two-pass softmax: (log2e * max, sum exp(v - max)) of this tile's columns per row,
                             // [pixel][n_tiles][2] (the maximum in the log2 domain: the kernels use ex2); nothing else is written
    const float* rowstat_in; // pass 2: v <- exp(v - M) / L with (M, L) combined from the row's n_tiles partials
    int rowstat_nt;          // n_tiles of the pass-1 launch
    const float* rowsub;     // v <- (v - rowsub[pixel]) * mulin[pixel, c]   (dS = P o (dP - rowsum(dO o O)))
    const act_t* mulin;
    int mulin_C;
    // ---- both modes (TMA-I/O and row-fusion kernels): transposed 16-bit copy of the main output (FWD: the raw value,
    //      BWD: dx) for channels [outT_c0, outT_c1): outT[(n * (outT_c1 - outT_c0) + c - outT_c0) * H * W + h * W + w].
    //      The attention backward consumes P^T, dS^T, theta^T and dO^T as K-major operands; emitting them here replaces
    //      four transpose passes (a warp's 32 lanes are 32 consecutive pixels: two full 32-byte sectors per store).
    act_t* outT;
    int outT_c0, outT_c1;
    // ---- epilogue, backward
    const act_t* saved;  // forward activation of the layer being differentiated
    int saved_C;
    // BN-affine gradient sums, deterministic: every (image, partial, channel) slot is written by exactly one warp,
    // statp[((n * statp_parts + part) * 2 + {0: sum_pix dpre, 1: sum_pix dpre * saved}) * statp_C + c]; a fixed-order
    // reduction over `part` follows (stat_reduce_kernel). part = tile index inside the image (128-pixel tiles: the four
    // epilogue warps are summed in shared memory first) or the 32-row quarter of a small image.
    float* statp;
    int statp_parts, statp_C;
    const act_t* addin;  // gradient arriving through the skip connection
    int addin_C, addin_climit, addin_pool;  // pool: sum the 2x2 block of a [NI,2H,2W,addin_C] map
    act_t* dx;
    int dx_C;
    float* dx_f32;  // optional fp32 copy (used for the latent-side tensors)
    int dx_f32_C;
    // ---- StyleGAN2 modulated-convolution epilogues (FLAVOR_SG instantiations, sg_epilogue.cuh)
    const float* sg_dm;      // [NI, sg_ld] demodulation: of this layer (FWD) / of the layer that produced `saved` (BWD)
    int sg_ld;
    const float* sg_noise;   // [NI, H', W'] noise image at the OUTPUT resolution (FWD) / at this grid (BWD), or null
    const float* sg_nw;      // device scalar: noise strength
    const float* sg_bias;    // BWD: bias of the layer that produced `saved`
    int d2s_C;               // FWD: columns are (phase, channel), depth-to-space stores, d2s_C channels per phase (0: off)
    int s2d;                 // BWD: dx leaves in space-to-depth layout [NI, H/2, W/2, 4 * dx_C]
    act_t* dx2;              // BWD: optional copy of the gradient wrt `saved` before its activation's derivative
    // ---- tile order: walk the tiles from the last to the first. Consecutive layers alternate the direction
    //      ("serpentine"): a layer starts on the tiles its producer wrote LAST, which are still in the 126 MB L2.
    int tile_reverse;
};

// DEEP: one CTA per SM with the full shared memory as pipeline (8 x 24 KB stages for BN = 64): for launches with
// fewer tiles than SMs and a long K loop (the 4x4 .. 16x16 generator blocks: 24-144 tiles, K up to 4608), where
// the main loop is bound by the bytes ONE CTA keeps in flight, not by the tensor pipe.
template <int BN, bool TMA_OUT = false, bool DEEP = false>
struct GemmCfg {
    static constexpr int kBTileBytes = BN * kBK * 2;
    static constexpr int kStageBytes = kATileBytes + kBTileBytes;
    // Two co-resident CTAs per SM for BN <= 128: the epilogue (global loads/stores issued by only
    // four warps) is latency-bound, a second CTA doubles the bytes in flight and lets one CTA's
    // epilogue overlap the other's main loop even on single-tile launches.
    // TMA_OUT ("TMA I/O" variant, used where the epilogue dominates): one CTA per SM with a fifth
    // role, the epilogue-input loader, and TWO epilogue groups of four warps: group g drains TMEM
    // accumulator stage g (tiles alternate between the groups), so two tiles' epilogues run
    // concurrently and each scheduler has two epilogue warps to hide latency with. Outputs leave
    // through 16 KB swizzled slabs (128 rows x 64 channels, one {raw, act} pair per group) and
    // cp.async.bulk.tensor stores; the per-pixel epilogue inputs (saved activation, residual skip,
    // skip gradient) are prefetched by TMA into a ring of kInSlots slabs shared by both groups, so
    // no thread ever waits on a global load.
    static constexpr int kOcc = (TMA_OUT || DEEP) ? 1 : ((BN <= 128) ? P2L_OCC : 1);
    static constexpr int kInSlots = TMA_OUT ? 4 : 0;
    static constexpr int kOutBytes = TMA_OUT ? (4 + kInSlots) * kATileBytes : 0;  // 2 x {raw, act} + inputs
    static constexpr int kEpiGroups = (TMA_OUT || kOcc == 1) ? 2 : 1;  // one CTA per SM: two groups drain the two TMEM stages
    static constexpr int kThreads = 64 + 128 * kEpiGroups + (TMA_OUT ? 32 * kEpiGroups : 0);  // + one input-loader warp per group
    static constexpr int kMaxStages = ((kOcc == 2 ? 104 : 208) * 1024 - kOutBytes) / kStageBytes;
    static constexpr int kStages = kMaxStages > 8 ? 8 : kMaxStages;
    static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + 1024 /*align slack*/ + 256 /*barriers*/ + 6 * BN * 4 /*coefficient tables*/ + kStatRedBytes;
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

// Row `row` (0..127) of a 128-row x 128-byte slab in the TMA 128B-swizzle layout: write 32 bf16
// (pieces piece0 .. piece0+3 of the row's eight 16-byte pieces).
__device__ __forceinline__ void slab_put_row(uint8_t* buf, int row, int piece0, const float (&v)[32]) {
    const uint32_t base = smem_u32(buf) + row * 128;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        sts128(base + (((piece0 + q) ^ (row & 7)) << 4),
               make_uint4(pack_bf16(v[q * 8], v[q * 8 + 1]), pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                          pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16(v[q * 8 + 6], v[q * 8 + 7])));
    }
}
__device__ __forceinline__ void slab_put_row(uint8_t*, int, int, const float (&)[16]) {}
// v[32] += the 32 bf16 of row `row`, pieces piece0 .. piece0+3, of a swizzled 128-byte-row slab
__device__ __forceinline__ void slab_add_row(const uint8_t* buf, int row, int piece0, float (&v)[32]) {
    const uint32_t base = smem_u32(buf) + row * 128;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 t = lds128(base + (((piece0 + q) ^ (row & 7)) << 4));
        v[q * 8 + 0] += bf16_lo(t.x); v[q * 8 + 1] += bf16_hi(t.x);
        v[q * 8 + 2] += bf16_lo(t.y); v[q * 8 + 3] += bf16_hi(t.y);
        v[q * 8 + 4] += bf16_lo(t.z); v[q * 8 + 5] += bf16_hi(t.z);
        v[q * 8 + 6] += bf16_lo(t.w); v[q * 8 + 7] += bf16_hi(t.w);
    }
}
__device__ __forceinline__ void slab_add_row(const uint8_t*, int, int, float (&)[16]) {}

// Row-per-thread global access helpers: CH 16-bit values (CH*2 bytes) of one pixel row. `wide`: the row
// pitch keeps every chunk 32-byte aligned, so 256-bit accesses are legal.
__device__ __forceinline__ bool wide_ok(const void* base, int row_pitch_bytes) {
    return ((reinterpret_cast<uintptr_t>(base) | static_cast<uintptr_t>(row_pitch_bytes)) & 31) == 0;
}
__device__ __forceinline__ uint4 pack8_act(const float* v) {
    return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ void acc8_act(const uint4 t, float* v) {
    v[0] += bf16_lo(t.x); v[1] += bf16_hi(t.x); v[2] += bf16_lo(t.y); v[3] += bf16_hi(t.y);
    v[4] += bf16_lo(t.z); v[5] += bf16_hi(t.z); v[6] += bf16_lo(t.w); v[7] += bf16_hi(t.w);
}
template <int CH>
__device__ __forceinline__ void row_store(act_t* dst, const float (&v)[CH], bool wide) {
    if (CH == 32 && wide) {
#pragma unroll
        for (int q = 0; q < CH / 16; ++q) stg256(dst + q * 16, pack8_act(v + q * 16), pack8_act(v + q * 16 + 8));
    } else {
#pragma unroll
        for (int q = 0; q < CH / 8; ++q) reinterpret_cast<uint4*>(dst)[q] = pack8_act(v + q * 8);
    }
}
template <int CH>
__device__ __forceinline__ void row_store_packed(act_t* dst, const uint4 (&o)[CH / 8], bool wide) {
    if (CH == 32 && wide) {
#pragma unroll
        for (int q = 0; q < CH / 16; ++q) stg256(dst + q * 16, o[2 * q], o[2 * q + 1]);
    } else {
#pragma unroll
        for (int q = 0; q < CH / 8; ++q) reinterpret_cast<uint4*>(dst)[q] = o[q];
    }
}
// v[0..CH) += row
template <int CH>
__device__ __forceinline__ void row_load_add(const act_t* src, float (&v)[CH], bool wide) {
    if (CH == 32 && wide) {
#pragma unroll
        for (int q = 0; q < CH / 16; ++q) {
            uint4 a, b;
            ldg256(src + q * 16, a, b);
            acc8_act(a, v + q * 16);
            acc8_act(b, v + q * 16 + 8);
        }
    } else {
#pragma unroll
        for (int q = 0; q < CH / 8; ++q) acc8_act(__ldg(reinterpret_cast<const uint4*>(src) + q), v + q * 8);
    }
}
template <int CH>
__device__ __forceinline__ void row_store_f32(float* dst, const float (&v)[CH], bool wide) {
    if (wide) {
#pragma unroll
        for (int q = 0; q < CH / 8; ++q)
            stg256(dst + q * 8, make_uint4(__float_as_uint(v[q * 8]), __float_as_uint(v[q * 8 + 1]), __float_as_uint(v[q * 8 + 2]), __float_as_uint(v[q * 8 + 3])),
                   make_uint4(__float_as_uint(v[q * 8 + 4]), __float_as_uint(v[q * 8 + 5]), __float_as_uint(v[q * 8 + 6]), __float_as_uint(v[q * 8 + 7])));
    } else {
#pragma unroll
        for (int q = 0; q < CH / 4; ++q) reinterpret_cast<float4*>(dst)[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
    }
}

// transposed 16-bit copy of a chunk of the thread's output row (ConvGemmParams::outT)
// L2 prefetch of NBYTES contiguous bytes (one request per 128-byte line)
template <int NBYTES>
__device__ __forceinline__ void prefetch_l2_row(const void* ptr) {
#pragma unroll
    for (int o = 0; o < NBYTES; o += 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const char*>(ptr) + o));
}
template <int CH>
__device__ __forceinline__ void row_store_transposed(const ConvGemmParams& p, int n, int h, int w, int cbase, const float (&v)[CH]) {
    const long hw = static_cast<long>(p.H) * p.W;
    act_t* dst = p.outT + (static_cast<long>(n) * (p.outT_c1 - p.outT_c0) + (cbase - p.outT_c0)) * hw + static_cast<long>(h) * p.W + w;
    if (cbase >= p.outT_c0 && cbase + CH <= p.outT_c1) {   // the whole chunk lies inside the slice: no per-element tests
#pragma unroll
        for (int j = 0; j < CH; ++j) dst[j * hw] = f2a(v[j]);
        return;
    }
#pragma unroll
    for (int j = 0; j < CH; ++j)
        if (cbase + j >= p.outT_c0 && cbase + j < p.outT_c1) dst[j * hw] = f2a(v[j]);
}

// tile index of the i-th work item (ConvGemmParams::tile_reverse)
__device__ __forceinline__ int tile_of(const ConvGemmParams& p, int i, int total_tiles) {
    return p.tile_reverse ? total_tiles - 1 - i : i;
}

// Direct epilogue: every thread reads / writes the global rows of its own accumulator row.
// ROWFUSE: the attention instantiations (row-wise softmax statistics / normalisation / softmax gradient); kept out of
// every other kernel — the epilogue body is the hot loop of 8 warps per SM and its size is what the instruction cache sees.
template <int BN, int MODE, int CH, bool TMA_OUT, int NG, int FLAVOR = FLAVOR_PLAIN>
__device__ __forceinline__ void epilogue_loop_direct(const ConvGemmParams& p, const CUtensorMap* tmo, uint8_t* obuf,
                                                     uint64_t* in_full, uint64_t* in_empty, uint64_t* tfull_bar,
                                                     uint64_t* tempty_bar, uint32_t tmem_base, int total_tiles, int warp,
                                                     int lane, float* ctab) {
    constexpr bool ROWFUSE = (FLAVOR == FLAVOR_ROWFUSE);
    // the 4 KB behind the tables: cross-warp staging of the BN-gradient column sums, [group][parity][quad][2][32]
    float* stat_red = ctab + 6 * BN;
    int stat_it = 0;
    // ctab: 2 x 3 x BN floats in shared memory — per-tile copies of bias / affine gain / affine offset
    // (they depend on (image, channel) only; staging them once per tile, before the accumulator is
    // ready, takes their global-load latency off the per-chunk critical path)
    const bool use_tab = (p.nb == 1);
    // TMA_OUT: tmo[0] = raw / dx, tmo[1] = act, tmo[2] = act on the 2x grid (5-D view), tmo[3] = act_lo;
    // obuf = [2 output slabs][kInSlots input slabs]; input slabs arrive through in_full / in_empty
    constexpr int NIN = GemmCfg<BN, TMA_OUT>::kInSlots;
    const int grp = (NG == 2) ? ((warp - 2) >> 2) : 0;  // epilogue group = TMEM accumulator stage it drains
    // the group's first warp issues the bulk stores: under elect.sync (deterministic leader, so the same thread
    // commits and waits on the bulk groups), never from `if (lane == 0)` (see conv_gemm_kernel)
    const bool store_warp = TMA_OUT && ((warp - 2) & 3) == 0;
    uint8_t* inbuf = nullptr;
    constexpr int RING = (NIN >= NG) ? NIN / NG : 1;
    if constexpr (TMA_OUT) {
        // each group has its own ring of RING input slots, filled by the loader warp in this group's
        // consumption order (an mbarrier parity wait is only safe for an in-order consumer)
        inbuf = obuf + (4 + grp * RING) * kATileBytes;
        in_full += grp * RING;
        in_empty += grp * RING;
        obuf += grp * 2 * kATileBytes;  // this group's {raw | dx, act} slab pair
    }
    // one output tensor only (dgrad, or act without raw): the two slabs of the pair double-buffer it
    const bool single_out = (MODE == EPI_BWD) || (p.raw == nullptr);
    int slab_no = 0;
    uint8_t *s_raw = obuf, *s_act = obuf;
    int in_cnt = 0;               // input slabs this group has consumed
    int slot0 = 0, slot1 = 0;     // slots of the current 64-channel slab
    bool has0 = false, has1 = false;
    // ------------------------------------------------------------------ epilogue warps
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;
    const int wi = row % p.tw;
    const int hi = (row / p.tw) % p.th;
    const int ni = row / (p.tw * p.th);
    const int rows_per_img = p.tw * p.th;
    float alpha = p.alpha;
    if (p.alpha_ptr) alpha *= __ldg(p.alpha_ptr);
    int it = grp;  // index of the tile in this CTA's sequence
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
        const long pix = (static_cast<long>(n) * p.H + h) * p.W + w;
        if constexpr (MODE == EPI_BWD && !TMA_OUT) {
            // this thread's saved-activation row on its way into L2 while the tile's main loop still runs: the row load of
            // the first chunk otherwise waits a full HBM round trip after the accumulator is ready
            if (p.saved && valid) prefetch_l2_row<BN * 2>(p.saved + pix * p.saved_C + n_tile * BN);
        }

        const uint32_t tab = smem_u32(ctab) + (it & 1) * 3 * BN * 4;  // shared-space byte address
        if (use_tab) {
            // two groups without the slab barriers of the TMA path: this group's previous tile used the
            // same table half, so every warp must have finished reading it
            if constexpr (NG == 2 && !TMA_OUT) bar_epilogue(grp);
            const int et = ((warp - 2) & 3) * 32 + lane;  // 0..127 within the group
            const int nn = min(tni * p.nb, p.NI - 1);
            for (int j = et; j < BN; j += 128) {
                const int ch = n_tile * BN + j;
                const bool in = ch < p.Cout;
                sts32f(tab + j * 4, (MODE == EPI_FWD && p.bias && in) ? __ldg(p.bias + ch) : 0.f);
                sts32f(tab + (BN + j) * 4, (p.aff_a && in) ? __ldg(p.aff_a + static_cast<long>(nn) * p.aff_stride + ch) : 1.f);
                sts32f(tab + (2 * BN + j) * 4, (MODE == EPI_FWD && p.aff_s && in) ? __ldg(p.aff_s + static_cast<long>(nn) * p.aff_stride + ch) : 0.f);
            }
            bar_epilogue(grp);  // table[it & 1] was last read by this group's previous tile (or two tiles ago): every warp has passed a barrier since
        }
        {
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
        }
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
        // row-wise softmax fusions (forward mode, direct path): running (max, sum) of pass 1 / (M, 1/L) of pass 2
        [[maybe_unused]] float rs_max = -INFINITY, rs_sum = 0.f, rs_M = 0.f, rs_invL = 1.f, rs_sub = 0.f;
        if constexpr (ROWFUSE && MODE == EPI_FWD && !TMA_OUT) {
            if (p.rowstat_in && valid) {
                const float* rp = p.rowstat_in + pix * p.rowstat_nt * 2;
                float M = -INFINITY;
                for (int t = 0; t < p.rowstat_nt; ++t) M = fmaxf(M, __ldg(rp + 2 * t));
                float L = 0.f;
                for (int t = 0; t < p.rowstat_nt; ++t) L += __ldg(rp + 2 * t + 1) * exp2f(__ldg(rp + 2 * t) - M);   // maxima in the log2 domain
                rs_M = M;
                rs_invL = 1.f / L;
            }
            if (p.rowsub && valid) rs_sub = __ldg(p.rowsub + pix);
        }

#pragma unroll 1
        for (int c = 0; c < BN; c += CH) {
            const int cbase = n_tile * BN + c;
            if (cbase >= p.Cout) break;  // warp-uniform
            if constexpr (TMA_OUT) {
                if ((c & 63) == 0) {
                    // both output slabs are about to be overwritten: their previous stores must have
                    // finished READING shared memory
                    // the pair written two slabs ago must have been READ by its bulk stores
                    if (store_warp) {
                        if (elect_one()) {
                            if (single_out) bulk_wait_read1();
                            else bulk_wait_read0();
                        }
                        __syncwarp();
                    }
                    s_raw = single_out ? obuf + (slab_no & 1) * kATileBytes : obuf;
                    s_act = single_out ? s_raw : obuf + kATileBytes;
                    ++slab_no;
                    bar_epilogue(grp);
                    has0 = (MODE == EPI_FWD) ? (p.resid != nullptr) : (p.saved != nullptr);
                    has1 = (MODE == EPI_BWD) && p.addin != nullptr && cbase < p.addin_climit;
                    if (has0) {
                        slot0 = in_cnt % RING;
                        mbar_wait(&in_full[slot0], (in_cnt / RING) & 1);
                        ++in_cnt;
                    }
                    if (has1) {
                        slot1 = in_cnt % RING;
                        mbar_wait(&in_full[slot1], (in_cnt / RING) & 1);
                        ++in_cnt;
                    }
                }
            }
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
                if (use_tab) {
#pragma unroll
                    for (int q = 0; q < CH / 4; ++q) {
                        const float4 b4 = lds128f(tab + (c + q * 4) * 4);
                        v[q * 4 + 0] = fmaf(alpha, v[q * 4 + 0], b4.x); v[q * 4 + 1] = fmaf(alpha, v[q * 4 + 1], b4.y);
                        v[q * 4 + 2] = fmaf(alpha, v[q * 4 + 2], b4.z); v[q * 4 + 3] = fmaf(alpha, v[q * 4 + 3], b4.w);
                    }
                } else if (p.bias && full_chunk) {
#pragma unroll
                    for (int q = 0; q < CH / 4; ++q) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cbase) + q);
                        v[q * 4 + 0] = fmaf(alpha, v[q * 4 + 0], b4.x); v[q * 4 + 1] = fmaf(alpha, v[q * 4 + 1], b4.y);
                        v[q * 4 + 2] = fmaf(alpha, v[q * 4 + 2], b4.z); v[q * 4 + 3] = fmaf(alpha, v[q * 4 + 3], b4.w);
                    }
                } else if (!ROWFUSE || p.bias || alpha != 1.f) {   // the attention GEMMs have neither bias nor gain
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        float b = 0.f;
                        if (p.bias && cbase + j < p.Cout) b = __ldg(p.bias + cbase + j);
                        v[j] = alpha * v[j] + b;
                    }
                }
                if constexpr (TMA_OUT) {
                    if (has0) {
                        // residual skip from its prefetched slab (low-res box when the skip is upsampled)
                        const int sh = p.resid_shift;
                        const int rr = ((ni * (p.th >> sh)) + (hi >> sh)) * (p.tw >> sh) + (wi >> sh);
                        slab_add_row(inbuf + slot0 * kATileBytes, rr, (c & 32) >> 3, v);
                    }
                } else if (p.resid && valid) {
                    const int Hs = p.H >> p.resid_shift, Ws = p.W >> p.resid_shift;
                    const long rp = (static_cast<long>(n) * Hs + (h >> p.resid_shift)) * Ws + (w >> p.resid_shift);
                    row_load_add<CH>(p.resid + rp * p.resid_C + cbase, v, wide_ok(p.resid, p.resid_C * 2));
                }
                if constexpr (ROWFUSE && !TMA_OUT) {
                    // exponentials as ex2(v * log2e - m * log2e): one FFMA + one MUFU per element; the running maximum is
                    // kept in the log2 domain (rs_max = log2e * max)
                    constexpr float kLog2e = 1.4426950408889634f;
                    if (p.rowstat) {  // pass 1: online (max, sum exp) over this tile's columns; no stores
                        float cm = -INFINITY;
                        if (full_chunk) {
#pragma unroll
                            for (int j = 0; j < CH; ++j) cm = fmaxf(cm, v[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < CH; ++j) cm = (cbase + j < p.Cout) ? fmaxf(cm, v[j]) : cm;
                        }
                        const float nm = fmaxf(rs_max, cm * kLog2e);
                        float acc = 0.f;
                        if (full_chunk) {
#pragma unroll
                            for (int j = 0; j < CH; ++j) acc += exp2f(fmaf(v[j], kLog2e, -nm));
                        } else {
#pragma unroll
                            for (int j = 0; j < CH; ++j) acc += (cbase + j < p.Cout) ? exp2f(fmaf(v[j], kLog2e, -nm)) : 0.f;
                        }
                        rs_sum = rs_sum * exp2f(rs_max - nm) + acc;
                        rs_max = nm;
                        continue;
                    }
                    if (p.rowstat_in) {
#pragma unroll
                        for (int j = 0; j < CH; ++j) v[j] = exp2f(fmaf(v[j], kLog2e, -rs_M)) * rs_invL;
                    }
                    if (p.mulin) {
                        float mv[CH];
#pragma unroll
                        for (int j = 0; j < CH; ++j) mv[j] = 0.f;
                        if (valid) row_load_add<CH>(p.mulin + pix * p.mulin_C + cbase, mv, wide_ok(p.mulin, p.mulin_C * 2));
#pragma unroll
                        for (int j = 0; j < CH; ++j) v[j] = (v[j] - rs_sub) * mv[j];
                    }
                }
                if (p.img_nchw) {
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < CH; ++j) {
                            if (cbase + j < p.Cout) {
                                p.img_nchw[((static_cast<long>(n) * p.Cout + cbase + j) * p.H + h) * p.W + w] = p.img_linear ? v[j] : tanhf(v[j]);
                            }
                        }
                    }
                }
                if (p.raw_f32 && valid) {
                    row_store_f32<CH>(p.raw_f32 + pix * p.raw_f32_C + cbase, v, wide_ok(p.raw_f32, p.raw_f32_C * 4));
                }
                if constexpr (TMA_OUT || ROWFUSE) {
                    if (p.outT && valid && cbase < p.outT_c1 && cbase + CH > p.outT_c0) row_store_transposed<CH>(p, n, h, w, cbase, v);
                }
                if constexpr (TMA_OUT) {
                    if (p.raw) slab_put_row(s_raw, row, (c & 32) >> 3, v);
                } else if (p.raw && valid) {
                    row_store<CH>(p.raw + pix * p.raw_C + cbase, v, wide_ok(p.raw, p.raw_C * 2));
                }
                if (p.act) {
                    if (p.aff_a && use_tab) {
#pragma unroll
                        for (int q = 0; q < CH / 4; ++q) {
                            const float4 a4 = lds128f(tab + (BN + c + q * 4) * 4);
                            const float4 s4 = lds128f(tab + (2 * BN + c + q * 4) * 4);
                            v[q * 4 + 0] = fmaf(a4.x, v[q * 4 + 0], s4.x);
                            v[q * 4 + 1] = fmaf(a4.y, v[q * 4 + 1], s4.y);
                            v[q * 4 + 2] = fmaf(a4.z, v[q * 4 + 2], s4.z);
                            v[q * 4 + 3] = fmaf(a4.w, v[q * 4 + 3], s4.w);
                        }
                    } else if (p.aff_a) {
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
                    if constexpr (TMA_OUT) {
                        slab_put_row(s_act, row, (c & 32) >> 3, v);
                    } else if (valid) {
                        uint4 o[CH / 8];
#pragma unroll
                        for (int q = 0; q < CH / 8; ++q) {
                            o[q] = make_uint4(pack_bf16(v[q * 8], v[q * 8 + 1]), pack_bf16(v[q * 8 + 2], v[q * 8 + 3]),
                                              pack_bf16(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16(v[q * 8 + 6], v[q * 8 + 7]));
                        }
                        const bool wide = wide_ok(p.act, p.act_C * 2) && (!p.act_lo || wide_ok(p.act_lo, 32));
                        if (!p.act_up) {
                            row_store_packed<CH>(p.act + pix * p.act_C + cbase, o, wide);
                        } else {
                            const int W2 = p.W * 2, H2 = p.H * 2;
#pragma unroll
                            for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
                                for (int dxx = 0; dxx < 2; ++dxx) {
                                    const long hp = (static_cast<long>(n) * H2 + 2 * h + dy) * W2 + 2 * w + dxx;
                                    row_store_packed<CH>(p.act + hp * p.act_C + cbase, o, wide);
                                }
                            }
                            if (p.act_lo) row_store_packed<CH>(p.act_lo + pix * p.act_C + cbase, o, wide);
                        }
                    }
                }
            } else {
                // ---------------------------------------------------------- backward
                float y[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j) { y[j] = 0.f; v[j] *= alpha; }
                if (TMA_OUT && p.saved) {
                    slab_add_row(inbuf + slot0 * kATileBytes, row, (c & 32) >> 3, y);  // y starts at 0
#pragma unroll
                    for (int j = 0; j < CH; ++j) v[j] = (valid && y[j] > 0.f) ? v[j] : 0.f;
                } else if (p.saved) {
                    if (valid) {
                        row_load_add<CH>(p.saved + pix * p.saved_C + cbase, y, wide_ok(p.saved, p.saved_C * 2));  // y starts at 0
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
                if (p.statp) {
                    // BN-affine gradients: reduce over the pixels (lanes) of one image; no atomics — every slot of the
                    // partial buffer has exactly one writer, so the step is bitwise reproducible
                    static_assert(CH == 32 || CH == 16, "chunk");
                    if constexpr (CH == 32) {
                        if (rows_per_img >= 32) {
                            float t0[32], t1[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) { t0[j] = v[j]; t1[j] = v[j] * y[j]; }
                            float s0 = colsum_group<32>(t0, lane);
                            float s1 = colsum_group<32>(t1, lane);
                            if (p.nb == 1) {
                                // the group's four warps hold the four 32-row quarters of ONE image's tile: sum them in
                                // shared memory in a fixed order (double-buffered by chunk parity: one barrier per chunk)
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
                                // small images (32 or 64 pixels): the warp's 32 rows are one quarter / half of one image
                                const int nn = tni * p.nb + (quad * 32) / rows_per_img;
                                const int part = ((quad * 32) % rows_per_img) >> 5;
                                if (nn < p.NI && cbase + lane < p.Cout) {
                                    float* dst = p.statp + (static_cast<long>(nn) * p.statp_parts + part) * 2 * p.statp_C + cbase + lane;
                                    dst[0] = s0;
                                    dst[p.statp_C] = s1;
                                }
                            }
                        } else {
                            // 16 pixels per image (4x4 maps): half-warp groups, one partial per image
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                float t0[16], t1[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j) { t0[j] = v[half * 16 + j]; t1[j] = v[half * 16 + j] * y[half * 16 + j]; }
                                const float s0 = colsum_group<16>(t0, lane);
                                const float s1 = colsum_group<16>(t1, lane);
                                const int nn = tni * p.nb + (quad * 32 + (lane & 16)) / rows_per_img;
                                const int ch = cbase + half * 16 + (lane & 15);
                                if (nn < p.NI && ch < p.Cout) {
                                    float* dst = p.statp + static_cast<long>(nn) * p.statp_parts * 2 * p.statp_C + ch;
                                    dst[0] = s0;
                                    dst[p.statp_C] = s1;
                                }
                            }
                        }
                    }
                }
                if (p.aff_a && use_tab) {
#pragma unroll
                    for (int q = 0; q < CH / 4; ++q) {
                        const float4 a4 = lds128f(tab + (BN + c + q * 4) * 4);
                        v[q * 4 + 0] *= a4.x; v[q * 4 + 1] *= a4.y; v[q * 4 + 2] *= a4.z; v[q * 4 + 3] *= a4.w;
                    }
                } else if (p.aff_a) {
                    const int nn = valid ? n : 0;
                    const float* pa = p.aff_a + static_cast<long>(nn) * p.aff_stride + cbase;
#pragma unroll
                    for (int q = 0; q < CH / 4; ++q) {
                        const float4 a4 = __ldg(reinterpret_cast<const float4*>(pa) + q);
                        v[q * 4 + 0] *= a4.x; v[q * 4 + 1] *= a4.y; v[q * 4 + 2] *= a4.z; v[q * 4 + 3] *= a4.w;
                    }
                }
                if constexpr (TMA_OUT) {
                    if (has1) slab_add_row(inbuf + slot1 * kATileBytes, row, (c & 32) >> 3, v);
                } else if (p.addin && valid && cbase < p.addin_climit) {
                    if (!p.addin_pool) {
                        row_load_add<CH>(p.addin + pix * p.addin_C + cbase, v, wide_ok(p.addin, p.addin_C * 2));
                    } else {
                        const int W2 = p.W * 2, H2 = p.H * 2;
#pragma unroll
                        for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
                            for (int dxx = 0; dxx < 2; ++dxx) {
                                const long hp = (static_cast<long>(n) * H2 + 2 * h + dy) * W2 + 2 * w + dxx;
                                row_load_add<CH>(p.addin + hp * p.addin_C + cbase, v, wide_ok(p.addin, p.addin_C * 2));
                            }
                        }
                    }
                }
                if constexpr (TMA_OUT) {
                    if (p.dx) slab_put_row(s_raw, row, (c & 32) >> 3, v);
                }
                if constexpr (TMA_OUT) {
                    if (p.outT && valid && cbase < p.outT_c1 && cbase + CH > p.outT_c0) row_store_transposed<CH>(p, n, h, w, cbase, v);
                }
                if (valid) {
                    if (!TMA_OUT && p.dx) {
                        row_store<CH>(p.dx + pix * p.dx_C + cbase, v, wide_ok(p.dx, p.dx_C * 2));
                    }
                    if (p.dx_f32) {
                        row_store_f32<CH>(p.dx_f32 + pix * p.dx_f32_C + cbase, v, wide_ok(p.dx_f32, p.dx_f32_C * 4));
                    }
                }
            }
            if constexpr (TMA_OUT) {
                if (c & 32) {  // a 64-channel slab is complete: hand it to the TMA store engine
                    if (has0 || has1) {
                        __syncwarp();  // every lane of this warp has read its rows of the input slabs
                        if (lane == 0) {
                            if (has0) mbar_arrive(&in_empty[slot0]);
                            if (has1) mbar_arrive(&in_empty[slot1]);
                        }
                    }
                    fence_async_smem();
                    bar_epilogue(grp);
                    if (store_warp && elect_one()) {
                        const int cs = n_tile * BN + (c & ~63);
                        const int w0 = twi * p.tw, h0 = thi * p.th, n0 = tni * p.nb;
                        if constexpr (MODE == EPI_FWD) {
                            if (p.raw) tma_store_4d(&tmo[0], s_raw, cs, w0, h0, n0);
                            if (p.act) {
                                if (!p.act_up) {
                                    tma_store_4d(&tmo[1], s_act, cs, w0, h0, n0);
                                } else {
                                    // nearest x2: the same slab goes to the four (dy, dx) phases of the 2x grid
                                    for (int d = 0; d < 4; ++d)
                                        tma_store_5d(&tmo[2], s_act, cs, d & 1, w0, d >> 1, n0 * p.H + h0);
                                    if (p.act_lo) tma_store_4d(&tmo[3], s_act, cs, w0, h0, n0);
                                }
                            }
                        } else {
                            if (p.dx) tma_store_4d(&tmo[0], s_raw, cs, w0, h0, n0);
                        }
                        bulk_commit();
                    }
                }
            }
        }
        if constexpr (ROWFUSE && MODE == EPI_FWD && !TMA_OUT) {
            if (p.rowstat && valid) {
                float* rp = p.rowstat + (pix * p.rowstat_nt + n_tile) * 2;
                rp[0] = rs_max;
                rp[1] = rs_sum;
            }
        }
        // all TMEM reads of this accumulator stage are complete (tmem_ld_wait above)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
    if (store_warp && elect_one()) bulk_wait0();  // shared memory must outlive the last bulk store's reads
}

}  // namespace p2l
#include "sg_epilogue.cuh"
namespace p2l {

struct OutMaps { CUtensorMap m[6]; };  // [0..3] epilogue outputs, [4..5] epilogue inputs

template <int BN, int MODE, bool TMA_OUT, bool DEEP = false, int FLAVOR = FLAVOR_PLAIN>
__global__ void __launch_bounds__(GemmCfg<BN, TMA_OUT, DEEP>::kThreads, GemmCfg<BN, TMA_OUT, DEEP>::kOcc)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ OutMaps tmO, const ConvGemmParams p) {
    using Cfg = GemmCfg<BN, TMA_OUT, DEEP>;
    constexpr int S = Cfg::kStages;
    constexpr int CH = (BN >= 32) ? 32 : 16;  // epilogue column chunk

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* obuf = smem + S * Cfg::kStageBytes;  // TMA_OUT: two 16 KB output slabs (1024-aligned)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes + Cfg::kOutBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + S;
    uint64_t* tfull_bar = bars + 2 * S;
    uint64_t* tempty_bar = bars + 2 * S + 2;
    constexpr int NIN = Cfg::kInSlots;
    uint64_t* in_full = bars + 2 * S + 4;
    uint64_t* in_empty = in_full + NIN;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4 + 2 * NIN);

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
        for (int s = 0; s < NIN; ++s) {
            mbar_init(&in_full[s], 1);
            mbar_init(&in_empty[s], 4);
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
    // Programmatic dependent launch ("pdl" option): everything above (barrier init, TMEM allocation, descriptor
    // prefetch) touches no global memory and may overlap the tail of the previous kernel in the stream; from here
    // on the previous grid's writes are needed. No-ops when the launch carries no PDL attribute.
    grid_dep_wait();
    grid_dep_launch_dependents();

    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int total_tiles = m_tiles * p.n_tiles;
    const int k_blocks = p.taps_h * p.taps_w * p.cin_chunks;
    const int Cin = p.cin_chunks * kBK;

    // The producer and MMA warps run their loops CONVERGED (all 32 lanes wait on the barriers and keep the loop
    // state); only the asynchronous-instruction issue itself sits under elect.sync. Issuing from `if (lane == 0)`
    // makes the compiler wrap every uniform-datapath instruction (UTCHMMA / UTMALDG / UTCBAR) in an
    // ELECT + BRA.U.ANY "waterfall" loop — measured (ncu source page, profiles/r1i): the MMA warp then spends ~75 %
    // of its time in issue overhead, ~160-195 cycles per MMA instruction.
    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        int stage = 0;
        uint32_t phase = 0;
        for (int wt = blockIdx.x; wt < total_tiles; wt += gridDim.x) {
            const int tile = tile_of(p, wt, total_tiles);
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
                        if (elect_one()) {
                            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                            tma_load_4d(sA, &tmA, &full_bar[stage], p.a_c0 + cc * kBK,
                                        w0 + s - p.pad_w, h0 + r - p.pad_h, n0);
                            tma_load_3d(sB, &tmB, &full_bar[stage], kbase + cc * kBK, n_tile * BN,
                                        p.b_batched ? n0 : 0);
                        }
                        __syncwarp();
                        if (++stage == S) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
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
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        // advance 16 bf16 = 32 B along K inside the swizzle row: +2 in 16-B units
                        umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == S) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (elect_one()) umma_commit(&tfull_bar[as]);
            __syncwarp();
        }
    } else
    if (warp >= 2 + 4 * Cfg::kEpiGroups) {
        // ------------------------------------------------------------------ epilogue-input loaders (one warp per group:
        // a single in-order loader would stall group 1's ring behind group 0's full one)
        if constexpr (TMA_OUT) {
            uint8_t* inbuf = obuf + 4 * kATileBytes;
            const bool has0 = (MODE == EPI_FWD) ? (p.resid != nullptr) : (p.saved != nullptr);
            const int sh = (MODE == EPI_FWD) ? p.resid_shift : 0;
            const uint32_t bytes0 = kATileBytes >> (2 * sh);
            constexpr int RING = NIN / Cfg::kEpiGroups;  // one ring per epilogue group
            const int g = warp - (2 + 4 * Cfg::kEpiGroups);
            int cnt = 0;
            for (int wt = blockIdx.x + g * gridDim.x; wt < total_tiles; wt += Cfg::kEpiGroups * gridDim.x) {
                const int tile = tile_of(p, wt, total_tiles);
                const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
                const int w0 = (m_tile % p.tiles_w) * p.tw;
                const int h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.th;
                const int n0 = (m_tile / (p.tiles_w * p.tiles_h)) * p.nb;
                for (int c = 0; c < BN; c += 64) {
                    const int cs = n_tile * BN + c;
                    if (cs >= p.Cout) break;
                    if (has0) {
                        const int slot = g * RING + cnt % RING;
                        mbar_wait(&in_empty[slot], ((cnt / RING) & 1) ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(&in_full[slot], bytes0);
                            tma_load_4d(inbuf + slot * kATileBytes, &tmO.m[4], &in_full[slot], cs, w0 >> sh, h0 >> sh, n0);
                        }
                        __syncwarp();
                        ++cnt;
                    }
                    if (MODE == EPI_BWD && p.addin != nullptr && cs < p.addin_climit) {
                        const int slot = g * RING + cnt % RING;
                        mbar_wait(&in_empty[slot], ((cnt / RING) & 1) ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(&in_full[slot], kATileBytes);
                            tma_load_4d(inbuf + slot * kATileBytes, &tmO.m[5], &in_full[slot], cs, w0, h0, n0);
                        }
                        __syncwarp();
                        ++cnt;
                    }
                }
            }
        }
    } else {
        float* ctab = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes + Cfg::kOutBytes + 256);
        if constexpr (FLAVOR == FLAVOR_SG) {
            epilogue_loop_sg<BN, MODE, Cfg::kEpiGroups>(p, tfull_bar, tempty_bar, tmem_base, total_tiles, warp, lane, ctab);
        } else {
            epilogue_loop_direct<BN, MODE, CH, TMA_OUT, Cfg::kEpiGroups, FLAVOR>(p, tmO.m, obuf, in_full, in_empty, tfull_bar, tempty_bar,
                                                                                tmem_base, total_tiles, warp, lane, ctab);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// ============================================================================= 3x3, halo patches
// For 3x3 / pad 1 convolutions the plain kernel above fetches the A tile once per filter tap:
// nine 16 KB loads of almost the same pixels, and with small N (64..128 output channels) the
// layer is bound by L2 -> SM bandwidth, not by the tensor pipe. Here the producer loads ONE
// (16+2) x (8+2) pixel patch per 64-channel chunk (box {64, P, 18, 1}, P = patch row pitch in
// pixels) and the nine taps are nine views of that patch: tap (r, s) starts (r*P + s) rows in,
// 8-row groups (one output row of the 8-wide tile) are P*128 bytes apart (the descriptor's SBO).
// TMA and tcgen05 both apply the 128-byte swizzle on absolute shared-memory address bits, so a
// row-shifted start address addresses the same bytes TMA wrote.
// Two rings: A patches (consumed once per 9 taps) and B weight tiles (one per tap).
template <int BN, int P>
struct HaloCfg {
    static constexpr int kTw = 8, kTh = 16;
    static constexpr int kPatchTx = P * (kTh + 2) * 128;
    static constexpr int kPatchBytes = ((kPatchTx + 1023) / 1024) * 1024;
    static constexpr int kBTileBytes = BN * kBK * 2;
    static constexpr int kThreads = 64 + 2 * 128;   // TMA, MMA, two epilogue groups (one per TMEM stage)
    static constexpr int kMaxBars = 64;             // 8-byte slots reserved for barriers
    static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int kFixedBytes = 1024 /*align slack*/ + kMaxBars * 8 + 6 * BN * 4 /*coefficient tables*/ + kStatRedBytes;
    static constexpr int kMaxSmem = 227 * 1024;
};

// Weights: when the whole [BN x 9*Cin] matrix fits next to the patch ring (p.halo_resb; the 64-channel
// layers), it is loaded ONCE per persistent CTA and stays resident — per tile only the patch moves.
template <int BN, int MODE, int P, int FLAVOR = FLAVOR_PLAIN>
__global__ void __launch_bounds__(HaloCfg<BN, P>::kThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const ConvGemmParams p) {
    using Cfg = HaloCfg<BN, P>;
    constexpr int CH = (BN >= 32) ? 32 : 16;
    const int SA = p.halo_sa;
    const bool resb = p.halo_resb != 0;
    const int SB = resb ? 1 : p.halo_sb;                       // barrier pairs of the B ring
    const int nB = resb ? 9 * p.cin_chunks : p.halo_sb;        // B tiles held in shared memory
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smemB = smem + SA * Cfg::kPatchBytes;
    uint8_t* tail = smemB + nB * Cfg::kBTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    uint64_t* afull = bars;
    uint64_t* aempty = bars + SA;
    uint64_t* bfull = bars + 2 * SA;
    uint64_t* bempty = bars + 2 * SA + SB;
    uint64_t* tfull_bar = bars + 2 * SA + 2 * SB;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < SA; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
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
    const int Cin = p.cin_chunks * kBK;

    // producer / MMA warps converged, issue under elect.sync (see conv_gemm_kernel)
    if (warp == 0) {
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        if (resb && static_cast<int>(blockIdx.x) < total_tiles) {
            if (elect_one()) {
                mbar_expect_tx(&bfull[0], nB * Cfg::kBTileBytes);
                for (int tap = 0; tap < 9; ++tap)
                    for (int cc = 0; cc < p.cin_chunks; ++cc)
                        tma_load_3d(smemB + (tap * p.cin_chunks + cc) * Cfg::kBTileBytes, &tmB, &bfull[0], tap * Cin + cc * kBK, 0, 0);
            }
            __syncwarp();
        }
        for (int wt = blockIdx.x; wt < total_tiles; wt += gridDim.x) {
            const int tile = tile_of(p, wt, total_tiles);
            const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
            const int w0 = (m_tile % p.tiles_w) * Cfg::kTw;
            const int h0 = ((m_tile / p.tiles_w) % p.tiles_h) * Cfg::kTh;
            const int n0 = m_tile / (p.tiles_w * p.tiles_h);
            for (int cc = 0; cc < p.cin_chunks; ++cc) {
                mbar_wait(&aempty[sa], pa ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&afull[sa], Cfg::kPatchTx);
                    tma_load_4d(smem + sa * Cfg::kPatchBytes, &tmA, &afull[sa], p.a_c0 + cc * kBK, w0 - 1, h0 - 1, n0);
                }
                __syncwarp();
                if (++sa == SA) { sa = 0; pa ^= 1; }
                if (resb) continue;
                for (int tap = 0; tap < 9; ++tap) {
                    mbar_wait(&bempty[sb], pb ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&bfull[sb], Cfg::kBTileBytes);
                        tma_load_3d(smemB + sb * Cfg::kBTileBytes, &tmB, &bfull[sb], tap * Cin + cc * kBK, n_tile * BN, 0);
                    }
                    __syncwarp();
                    if (++sb == SB) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16(BN);
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        int it = 0;
        if (resb && static_cast<int>(blockIdx.x) < total_tiles) mbar_wait(&bfull[0], 0);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tempty_bar[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * BN;
            for (int cc = 0; cc < p.cin_chunks; ++cc) {
                mbar_wait(&afull[sa], pa);
                tc_fence_after();
                const uint32_t a_base = smem_u32(smem + sa * Cfg::kPatchBytes);
                if (resb) {
                    // weights resident: the nine taps of this channel chunk are 36 back-to-back MMAs
                    if (elect_one()) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint32_t b_addr = smem_u32(smemB + (tap * p.cin_chunks + cc) * Cfg::kBTileBytes);
                            const uint32_t a_addr = a_base + ((tap / 3) * P + (tap % 3)) * 128;
                            const uint64_t adesc = umma_desc_sw128(a_addr, P * 128, p.halo_bo ? (a_addr >> 7) : 0);
                            const uint64_t bdesc = umma_desc_k128(b_addr);
#pragma unroll
                            for (int k = 0; k < kBK / 16; ++k)
                                umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (cc | tap | k) != 0);
                        }
                        umma_commit(&aempty[sa]);
                    }
                    __syncwarp();
                } else {
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(&bfull[sb], pb);
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(smemB + sb * Cfg::kBTileBytes);
                        const int r = tap / 3, s = tap - 3 * r;
                        const uint32_t a_addr = a_base + (r * P + s) * 128;
                        const uint64_t adesc = umma_desc_sw128(a_addr, P * 128, p.halo_bo ? (a_addr >> 7) : 0);
                        const uint64_t bdesc = umma_desc_k128(b_addr);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < kBK / 16; ++k)
                                umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (cc | tap | k) != 0);
                            umma_commit(&bempty[sb]);
                        }
                        __syncwarp();
                        if (++sb == SB) { sb = 0; pb ^= 1; }
                    }
                    if (elect_one()) umma_commit(&aempty[sa]);
                    __syncwarp();
                }
                if (++sa == SA) { sa = 0; pa ^= 1; }
            }
            if (elect_one()) umma_commit(&tfull_bar[as]);
            __syncwarp();
        }
    } else {
        float* ctab = reinterpret_cast<float*>(tail + Cfg::kMaxBars * 8);
        if constexpr (FLAVOR == FLAVOR_SG) {
            epilogue_loop_sg<BN, MODE, 2>(p, tfull_bar, tempty_bar, tmem_base, total_tiles, warp, lane, ctab);
        } else {
            epilogue_loop_direct<BN, MODE, CH, false, 2>(p, nullptr, nullptr, nullptr, nullptr, tfull_bar, tempty_bar, tmem_base, total_tiles,
                                                         warp, lane, ctab);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

}  // namespace p2l
