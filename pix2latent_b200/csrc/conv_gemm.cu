// Host side of the tcgen05 implicit-GEMM: TMA descriptor construction, tile-shape choice,
// template dispatch. The kernel itself is in conv_gemm.cuh.
#include "conv_gemm.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <utility>
#include <vector>

namespace p2l {

// ----------------------------------------------------------------------------- errors
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

static long g_launches = 0;
void count_launch() { ++g_launches; }
void add_launches(long n) { g_launches += n; }
long launch_count() { return g_launches; }

static int g_tma_out = 1, g_tma_kmax = 512;
static int g_attn_fused = 1;  // "attn_fused": two-pass softmax + fused dS in the attention GEMM epilogues (no fp32 logits in HBM)
static int g_attn_emit_t = 1; // "attn_emit_t": the attention backward's transposed operands come out of the producing epilogues
static int g_serpentine = 1;  // "serpentine": consecutive layers walk their tiles in opposite directions (L2 reuse of the producer's last writes)
static int g_pdl = 1;  // "pdl": launch the tensor-core kernel with programmatic stream serialization (prologue overlaps the previous kernel's tail)
static int g_deep = 1, g_deep_kmin = 8;  // "deep": full-depth single-CTA pipeline for launches with <= #SM tiles and >= deep_kmin K blocks
static int g_grad_scale = (int)kGradScale;
float grad_scale() { return (float)g_grad_scale; }
// halo patches for 3x3 convs: "halo" = patch pitch in pixels (10 | 16), "halo_mode" = 0 off, 1 only where the
// weight matrix stays resident in shared memory (the 64-channel layers, L2-bandwidth-bound otherwise), 2 every
// eligible 3x3 conv
static int g_halo_rgb = 0;
static int g_halo = 10, g_halo_mode = 1, g_halo_bo = 0;  // mode 1: the 64-channel 3x3 layers (weights resident in shared memory) run on the halo-patch kernel: 121 -> 102 us at 256^2 (profiles/r1k)
void set_option(const char* key, int value) {
    if (!std::strcmp(key, "halo")) g_halo = value;
    else if (!std::strcmp(key, "grad_scale")) g_grad_scale = value > 0 ? value : 1;
    else if (!std::strcmp(key, "halo_bo")) g_halo_bo = value;
    else if (!std::strcmp(key, "halo_mode")) g_halo_mode = value;
    else if (!std::strcmp(key, "halo_rgb")) g_halo_rgb = value;
    else if (!std::strcmp(key, "tma_out")) g_tma_out = value;
    else if (!std::strcmp(key, "tma_kmax")) g_tma_kmax = value;
    else if (!std::strcmp(key, "pdl")) g_pdl = value;
    else if (!std::strcmp(key, "attn_fused")) g_attn_fused = value;
    else if (!std::strcmp(key, "attn_emit_t")) g_attn_emit_t = value;
    else if (!std::strcmp(key, "serpentine")) g_serpentine = value;
    else if (!std::strcmp(key, "deep")) g_deep = value;
    else if (!std::strcmp(key, "deep_kmin")) g_deep_kmin = value;
}
int get_option(const char* key) {
    if (!std::strcmp(key, "halo")) return g_halo;
    if (!std::strcmp(key, "grad_scale")) return g_grad_scale;
    if (!std::strcmp(key, "halo_bo")) return g_halo_bo;
    if (!std::strcmp(key, "halo_mode")) return g_halo_mode;
    if (!std::strcmp(key, "tma_out")) return g_tma_out;
    if (!std::strcmp(key, "tma_kmax")) return g_tma_kmax;
    if (!std::strcmp(key, "pdl")) return g_pdl;
    if (!std::strcmp(key, "attn_fused")) return g_attn_fused;
    if (!std::strcmp(key, "attn_emit_t")) return g_attn_emit_t;
    if (!std::strcmp(key, "serpentine")) return g_serpentine;
    if (!std::strcmp(key, "deep")) return g_deep;
    if (!std::strcmp(key, "deep_kmin")) return g_deep_kmin;
    return -1;
}

// ---- optional per-launch event timing of the tensor-core kernel (bench.py roofline leg)
static bool g_prof = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
static size_t g_prof_used = 0;
static double g_prof_flops = 0;
struct ProfRec { double flops; int BN, mode, halo, grid, M, N, K; };
static std::vector<ProfRec> g_prof_recs;
void profile_enable(int on) {
    g_prof = on != 0;
    g_prof_used = 0;
    g_prof_flops = 0;
    g_prof_recs.clear();
}
int profile_get(int i, float* ms, double* flops, int* info /*7 ints*/) {
    if (i < 0 || (size_t)i >= g_prof_used || (size_t)i >= g_prof_recs.size()) return -1;
    if (cudaEventSynchronize(g_prof_events[i].second) != cudaSuccess) return -1;
    cudaEventElapsedTime(ms, g_prof_events[i].first, g_prof_events[i].second);
    const ProfRec& r = g_prof_recs[i];
    *flops = r.flops;
    info[0] = r.BN; info[1] = r.mode; info[2] = r.halo; info[3] = r.grid; info[4] = r.M; info[5] = r.N; info[6] = r.K;
    return 0;
}
int profile_read(double* conv_ms, long* conv_launches, double* conv_flops) {
    double ms = 0;
    for (size_t i = 0; i < g_prof_used; ++i) {
        if (cudaEventSynchronize(g_prof_events[i].second) != cudaSuccess) { set_error("profile_read: event sync failed"); return -1; }
        float t = 0;
        cudaEventElapsedTime(&t, g_prof_events[i].first, g_prof_events[i].second);
        ms += t;
    }
    if (conv_ms) *conv_ms = ms;
    if (conv_launches) *conv_launches = (long)g_prof_used;
    if (conv_flops) *conv_flops = g_prof_flops;
    g_prof_used = 0;
    g_prof_flops = 0;
    return 0;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// ----------------------------------------------------------------------------- TMA maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess || !p) {
            set_error("cuTensorMapEncodeTiled not available from the driver");
            return nullptr;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int encode_bf16(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims,
                       const cuuint64_t* strides_bytes /*rank-1*/, const cuuint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return -1;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, P2L_TMAP_DTYPE, rank, const_cast<void*>(ptr), dims,
                    strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]",
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                  box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return -1;
    }
    return 0;
}

static int pow2_ceil(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

int conv_stat_parts_max(int H, int W) {
    int tw = pow2_ceil(W); if (tw > 16) tw = 16;
    int th = pow2_ceil(H); if (th > kBM / tw) th = kBM / tw;
    if (tw * th < kBM) return tw * th >= 32 ? (tw * th) / 32 : 1;
    const int a = ((W + tw - 1) / tw) * ((H + th - 1) / th);
    const int b = ((W + 7) / 8) * ((H + 15) / 16);   // halo-patch kernel tiling
    return a > b ? a : b;
}

// ----------------------------------------------------------------------------- build
int conv_op_build(ConvOp* op, const ConvDesc& d) {
    if (d.Cin <= 0 || d.Cin % kBK != 0) {
        set_error("conv_op_build: Cin=%d must be a positive multiple of %d", d.Cin, kBK);
        return -1;
    }
    if (d.A_C % 8 != 0 || (reinterpret_cast<uintptr_t>(d.A) & 15) || (reinterpret_cast<uintptr_t>(d.B) & 15)) {
        set_error("conv_op_build: A/B must be 16-byte aligned with channel count %% 8 == 0");
        return -1;
    }
    if (d.BN != 16 && d.BN != 64 && d.BN != 128 && d.BN != 256) {
        set_error("conv_op_build: unsupported BN=%d", d.BN);
        return -1;
    }
    std::memset(op, 0, sizeof(*op));
    ConvGemmParams p = d.epi;
    p.NI = d.NI; p.H = d.H; p.W = d.W;
    // tile box: up to 16 wide, up to 8 high, rest over images
    int tw = pow2_ceil(d.W); if (tw > 16) tw = 16;
    int th = pow2_ceil(d.H); if (th > kBM / tw) th = kBM / tw;
    int nb = kBM / (tw * th);
    // halo patches: always for the 3-channel rgb head (N = 16 tile: the layer is pure A traffic),
    // optional ("halo" option) elsewhere — measured neutral for N >= 64 (profiles/)
    const int halo_p = g_halo;
    const long resb_bytes = 9L * d.Cin * d.BN * 2;
    const bool resb = d.Cout <= d.BN && resb_bytes <= 80 * 1024;
    // mode 1: resident-weight layers on large images (measured: 256^2 gains 15-20 %, 128^2 loses 4 %). The choice depends
    // on the image size only, never on the batch: the tiling decides how the BN-gradient partial sums are grouped, and a
    // candidate's result must not depend on which other candidates share its launch (sub-batches, candidate sharding).
    const bool halo_on = (d.BN == 16) ? (g_halo_rgb != 0)
                                      : (g_halo_mode != 0 && (g_halo_mode == 2 || (resb && (long)d.H * d.W >= (1L << 16))));
    const bool halo = halo_p != 0 && halo_on && (!d.sg || (d.BN == 64 && halo_p == 10)) && d.kh == 3 && d.kw == 3 &&
                      d.pad_h == 1 && d.pad_w == 1 && d.B_batch == 0 && d.H >= 12 && d.W >= 8;
    if (halo) { tw = 8; th = 16; nb = 1; }
    if (d.B_batch > 0 && nb != 1) {
        set_error("conv_op_build: batched B needs >=128 pixels per image (got %dx%d)", d.H, d.W);
        return -1;
    }
    p.tw = tw; p.th = th; p.nb = nb;
    p.ltw = 0; while ((1 << p.ltw) < tw) ++p.ltw;
    p.lth = 0; while ((1 << p.lth) < th) ++p.lth;
    p.tiles_w = (d.W + tw - 1) / tw;
    p.tiles_h = (d.H + th - 1) / th;
    p.tiles_n = (d.NI + nb - 1) / nb;
    p.Cout = d.Cout;
    p.n_tiles = (d.Cout + d.BN - 1) / d.BN;
    p.taps_h = d.kh; p.taps_w = d.kw; p.pad_h = d.pad_h; p.pad_w = d.pad_w;
    p.cin_chunks = d.Cin / kBK;
    p.a_c0 = d.a_c0;
    p.b_batched = d.B_batch > 0 ? 1 : 0;
    p.halo_bo = g_halo_bo;
    p.halo_resb = (halo && resb) ? 1 : 0;
    p.halo_sa = 0; p.halo_sb = 0;
    if (halo) {
        // shared-memory plan: [patch ring][B tiles (resident matrix | ring)][barriers][coefficient tables]
        const long patch = (((long)halo_p * 18 * 128 + 1023) / 1024) * 1024, btile = (long)d.BN * kBK * 2;
        const long fixed = 1024 + 64 * 8 + 6L * d.BN * 4 + kStatRedBytes, budget = 227 * 1024 - fixed;
        if (p.halo_resb) {
            p.halo_sa = (int)((budget - resb_bytes) / patch);
        } else {
            p.halo_sb = d.BN == 256 ? 4 : (d.BN == 128 ? 6 : 8);
            p.halo_sa = (int)((budget - p.halo_sb * btile) / patch);
        }
        if (p.halo_sa > 6) p.halo_sa = 6;
        if (p.halo_sa < 2) { set_error("conv_op_build: halo plan does not fit shared memory"); return -1; }
        op->halo_smem = (int)(p.halo_sa * patch + (p.halo_resb ? resb_bytes : p.halo_sb * btile) + fixed);
    }
    if (p.alpha == 0.f) p.alpha = 1.f;
    // partial slots per image of the BN-gradient sums this launch fills (see ConvGemmParams::statp)
    op->stat_parts = (nb == 1) ? p.tiles_w * p.tiles_h : ((tw * th >= 32) ? (tw * th) / 32 : 1);
    if (d.epi.statp && op->stat_parts > d.epi.statp_parts) {
        set_error("conv_op_build: BN-gradient partial buffer holds %d slots per image, the launch needs %d", d.epi.statp_parts, op->stat_parts);
        return -1;
    }

    {   // A: {C, W, H, N}
        cuuint64_t dims[4] = {(cuuint64_t)d.A_C, (cuuint64_t)d.A_W, (cuuint64_t)d.A_H, (cuuint64_t)d.A_N};
        cuuint64_t str[3] = {(cuuint64_t)d.A_C * 2, (cuuint64_t)d.A_W * d.A_C * 2, (cuuint64_t)d.A_H * d.A_W * d.A_C * 2};
        cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)nb};
        if (halo) { box[1] = (cuuint32_t)halo_p; box[2] = (cuuint32_t)(th + 2); box[3] = 1; }
        if (encode_bf16(&op->tmA, d.A, 4, dims, str, box)) return -1;
    }
    {   // B: {K, Cout, batch}
        const cuuint64_t K = (cuuint64_t)d.kh * d.kw * d.Cin;
        cuuint64_t dims[3] = {K, (cuuint64_t)d.Cout, (cuuint64_t)(d.B_batch > 0 ? d.B_batch : 1)};
        cuuint64_t str[2] = {K * 2, K * 2 * (cuuint64_t)d.Cout};
        cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)d.BN, 1};
        if (encode_bf16(&op->tmB, d.B, 3, dims, str, box)) return -1;
    }
    // ---- epilogue outputs by TMA bulk store: worthwhile where the epilogue dominates (small K)
    const long Ktot = (long)d.kh * d.kw * d.Cin;
    const bool any_out = (d.epi.raw || d.epi.act || d.epi.dx) && !d.epi.rowstat && !d.epi.rowstat_in && !d.epi.mulin;  // row-wise fusions: direct epilogue only
    op->tma_out = (g_tma_out && !halo && !d.sg && any_out && (d.BN == 64 || d.BN == 128) && d.Cout % 64 == 0 && Ktot <= g_tma_kmax &&
                   !d.epi.img_nchw && !(d.epi.addin && d.epi.addin_pool) && (d.epi.resid_shift == 0 || (tw >= 2 && th >= 2)) &&
                   (!d.epi.addin || d.epi.addin_climit % 64 == 0)) ? 1 : 0;
    if (op->tma_out) {
        auto map4 = [&](CUtensorMap* m, const void* ptr, int C, int Hh, int Ww) -> int {
            cuuint64_t dims[4] = {(cuuint64_t)d.Cout, (cuuint64_t)Ww, (cuuint64_t)Hh, (cuuint64_t)d.NI};
            cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)Ww * C * 2, (cuuint64_t)Hh * Ww * C * 2};
            cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)nb};
            return encode_bf16(m, ptr, 4, dims, str, box);
        };
        // epilogue inputs: [4] = saved activation (bwd) / residual skip (fwd), [5] = skip gradient (bwd)
        auto map_in = [&](CUtensorMap* m, const void* ptr, int C, int Cext, int sh) -> int {
            const int Hh = d.H >> sh, Ww = d.W >> sh;
            cuuint64_t dims[4] = {(cuuint64_t)Cext, (cuuint64_t)Ww, (cuuint64_t)Hh, (cuuint64_t)d.NI};
            cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)Ww * C * 2, (cuuint64_t)Hh * Ww * C * 2};
            cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)(tw >> sh), (cuuint32_t)(th >> sh), (cuuint32_t)nb};
            return encode_bf16(m, ptr, 4, dims, str, box);
        };
        if (d.mode == EPI_FWD && d.epi.resid && map_in(&op->tmO.m[4], d.epi.resid, d.epi.resid_C, d.Cout, d.epi.resid_shift)) return -1;
        if (d.mode == EPI_BWD && d.epi.saved && map_in(&op->tmO.m[4], d.epi.saved, d.epi.saved_C, d.Cout, 0)) return -1;
        if (d.mode == EPI_BWD && d.epi.addin && map_in(&op->tmO.m[5], d.epi.addin, d.epi.addin_C, d.epi.addin_climit, 0)) return -1;
        const void* o0 = d.mode == EPI_FWD ? (const void*)d.epi.raw : (const void*)d.epi.dx;
        const int o0C = d.mode == EPI_FWD ? d.epi.raw_C : d.epi.dx_C;
        if (o0 && map4(&op->tmO.m[0], o0, o0C, d.H, d.W)) return -1;
        if (d.mode == EPI_FWD && d.epi.act) {
            if (!d.epi.act_up) {
                if (map4(&op->tmO.m[1], d.epi.act, d.epi.act_C, d.H, d.W)) return -1;
            } else {
                // [NI, 2H, 2W, C] viewed as {C, dx(2), W, dy(2), NI*H}
                const cuuint64_t C = (cuuint64_t)d.epi.act_C;
                cuuint64_t dims[5] = {(cuuint64_t)d.Cout, 2, (cuuint64_t)d.W, 2, (cuuint64_t)d.NI * d.H};
                cuuint64_t str[4] = {C * 2, 2 * C * 2, 2 * (cuuint64_t)d.W * C * 2, 4 * (cuuint64_t)d.W * C * 2};
                cuuint32_t box[5] = {(cuuint32_t)kBK, 1, (cuuint32_t)tw, 1, (cuuint32_t)(th * nb)};
                if (nb > 1 && th != d.H) { set_error("conv_op_build: act_up tile spans images with th != H"); return -1; }
                if (encode_bf16(&op->tmO.m[2], d.epi.act, 5, dims, str, box)) return -1;
                if (d.epi.act_lo && map4(&op->tmO.m[3], d.epi.act_lo, d.epi.act_C, d.H, d.W)) return -1;
            }
        }
    }
    const bool rowfuse = d.epi.rowstat || d.epi.rowstat_in || d.epi.mulin || d.epi.rowsub;
    if (rowfuse && (d.mode != EPI_FWD || halo || (d.BN != 64 && d.BN != 128))) {
        set_error("conv_op_build: row-wise softmax fusions run in the forward 1x1 kernels with BN 64 / 128");
        return -1;
    }
    if (d.epi.outT && !op->tma_out && !rowfuse) {
        set_error("conv_op_build: a transposed output needs the TMA-I/O or the row-fusion kernel (K <= tma_kmax, BN 64 / 128)");
        return -1;
    }
    if (d.sg && (rowfuse || op->tma_out || (d.BN != 64 && d.BN != 128 && d.BN != 256) || (halo && (d.BN != 64 || halo_p != 10)))) {
        set_error("conv_op_build: StyleGAN2 epilogues run in the direct-epilogue kernels with BN 64 / 128 / 256");
        return -1;
    }
    op->flavor = d.sg ? FLAVOR_SG : (rowfuse ? FLAVOR_ROWFUSE : FLAVOR_PLAIN);
    op->p = p;
    op->BN = d.BN;
    op->mode = d.mode;
    op->halo = halo ? halo_p : 0;
    const long total = (long)p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
    op->deep = (g_deep && !halo && !op->tma_out && !rowfuse && (d.BN == 64 || d.BN == 128) && total <= num_sms() &&
                (long)d.kh * d.kw * p.cin_chunks >= g_deep_kmin) ? 1 : 0;
    const long slots = (long)num_sms() * ((halo || d.BN > 128 || op->tma_out || op->deep) ? 1 : P2L_OCC);
    op->grid = (int)(total < slots ? total : slots);
    op->flops = 2.0 * d.NI * d.H * d.W * (double)d.Cout * d.kh * d.kw * d.Cin;
    return 0;
}

// ----------------------------------------------------------------------------- launch
template <int BN, int MODE, bool TMA_OUT, bool DEEP = false, int FLAVOR = FLAVOR_PLAIN>
static int launch_t(const ConvOp& op, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, TMA_OUT, DEEP>;
    static bool attr_set = false;
    if (!attr_set) {
        P2L_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_kernel<BN, MODE, TMA_OUT, DEEP, FLAVOR>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        attr_set = true;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (g_prof) {
        if (g_prof_used == g_prof_events.size()) {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            g_prof_events.emplace_back(e0, e1);
        }
        e0 = g_prof_events[g_prof_used].first;
        e1 = g_prof_events[g_prof_used].second;
        ++g_prof_used;
        g_prof_flops += op.flops;
        g_prof_recs.push_back({op.flops, op.BN, op.mode, op.halo, op.grid, op.p.NI * op.p.H * op.p.W, op.p.Cout,
                               op.p.taps_h * op.p.taps_w * op.p.cin_chunks * kBK});
        cudaEventRecord(e0, stream);
    }
    {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(op.grid);
        lc.blockDim = dim3(Cfg::kThreads);
        lc.dynamicSmemBytes = Cfg::kSmemBytes;
        lc.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at;
        lc.numAttrs = g_pdl ? 1 : 0;
        P2L_CUDA_CHECK(cudaLaunchKernelEx(&lc, conv_gemm_kernel<BN, MODE, TMA_OUT, DEEP, FLAVOR>, op.tmA, op.tmB, op.tmO, op.p));
    }
    if (g_prof) cudaEventRecord(e1, stream);
    count_launch();
    P2L_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int BN, int MODE, int P, int FLAVOR = FLAVOR_PLAIN>
static int launch_halo_t(const ConvOp& op, cudaStream_t stream) {
    using Cfg = HaloCfg<BN, P>;
    static bool attr_set = false;
    if (!attr_set) {
        P2L_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_halo_kernel<BN, MODE, P, FLAVOR>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kMaxSmem));
        attr_set = true;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (g_prof) {
        if (g_prof_used == g_prof_events.size()) {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            g_prof_events.emplace_back(e0, e1);
        }
        e0 = g_prof_events[g_prof_used].first;
        e1 = g_prof_events[g_prof_used].second;
        ++g_prof_used;
        g_prof_flops += op.flops;
        g_prof_recs.push_back({op.flops, op.BN, op.mode, op.halo, op.grid, op.p.NI * op.p.H * op.p.W, op.p.Cout,
                               op.p.taps_h * op.p.taps_w * op.p.cin_chunks * kBK});
        cudaEventRecord(e0, stream);
    }
    conv3x3_halo_kernel<BN, MODE, P, FLAVOR><<<op.grid, Cfg::kThreads, op.halo_smem, stream>>>(op.tmA, op.tmB, op.p);
    if (g_prof) cudaEventRecord(e1, stream);
    count_launch();
    P2L_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int conv_op_launch(const ConvOp& op, cudaStream_t stream) {
    if (op.flavor == FLAVOR_SG) {
        const bool f = op.mode == EPI_FWD;
        if (op.halo) {
            if (op.BN == 64 && op.halo == 10) return f ? launch_halo_t<64, EPI_FWD, 10, FLAVOR_SG>(op, stream) : launch_halo_t<64, EPI_BWD, 10, FLAVOR_SG>(op, stream);
            set_error("conv_op_launch: StyleGAN2 halo kernels exist for BN 64, pitch 10");
            return -1;
        }
        if (op.deep) {
            if (op.BN == 64) return f ? launch_t<64, EPI_FWD, false, true, FLAVOR_SG>(op, stream) : launch_t<64, EPI_BWD, false, true, FLAVOR_SG>(op, stream);
            if (op.BN == 128) return f ? launch_t<128, EPI_FWD, false, true, FLAVOR_SG>(op, stream) : launch_t<128, EPI_BWD, false, true, FLAVOR_SG>(op, stream);
        }
        if (op.BN == 64) return f ? launch_t<64, EPI_FWD, false, false, FLAVOR_SG>(op, stream) : launch_t<64, EPI_BWD, false, false, FLAVOR_SG>(op, stream);
        if (op.BN == 128) return f ? launch_t<128, EPI_FWD, false, false, FLAVOR_SG>(op, stream) : launch_t<128, EPI_BWD, false, false, FLAVOR_SG>(op, stream);
        if (op.BN == 256) return f ? launch_t<256, EPI_FWD, false, false, FLAVOR_SG>(op, stream) : launch_t<256, EPI_BWD, false, false, FLAVOR_SG>(op, stream);
        set_error("conv_op_launch: StyleGAN2 epilogues exist for BN 64 / 128 / 256 (got %d)", op.BN);
        return -1;
    }
    if (op.halo) {
#define P2L_HALO(bn, pp)                                                                   \
    if (op.BN == bn && op.halo == pp)                                                      \
        return op.mode == EPI_FWD ? launch_halo_t<bn, EPI_FWD, pp>(op, stream)             \
                                  : launch_halo_t<bn, EPI_BWD, pp>(op, stream);
        P2L_HALO(16, 10) P2L_HALO(64, 10) P2L_HALO(128, 10) P2L_HALO(256, 10)
        P2L_HALO(64, 16) P2L_HALO(128, 16) P2L_HALO(256, 16)
#undef P2L_HALO
        set_error("conv_op_launch: unsupported halo config BN=%d P=%d", op.BN, op.halo);
        return -1;
    }
    if (op.flavor == FLAVOR_ROWFUSE) {
        if (op.BN == 64) return launch_t<64, EPI_FWD, false, false, FLAVOR_ROWFUSE>(op, stream);
        if (op.BN == 128) return launch_t<128, EPI_FWD, false, false, FLAVOR_ROWFUSE>(op, stream);
        set_error("conv_op_launch: row-fusion epilogues exist for BN 64 / 128 (got %d)", op.BN);
        return -1;
    }
    if (op.tma_out) {
        if (op.BN == 64) return op.mode == EPI_FWD ? launch_t<64, EPI_FWD, true>(op, stream) : launch_t<64, EPI_BWD, true>(op, stream);
        if (op.BN == 128) return op.mode == EPI_FWD ? launch_t<128, EPI_FWD, true>(op, stream) : launch_t<128, EPI_BWD, true>(op, stream);
    }
    if (op.deep) {
        if (op.BN == 64) return op.mode == EPI_FWD ? launch_t<64, EPI_FWD, false, true>(op, stream) : launch_t<64, EPI_BWD, false, true>(op, stream);
        if (op.BN == 128) return op.mode == EPI_FWD ? launch_t<128, EPI_FWD, false, true>(op, stream) : launch_t<128, EPI_BWD, false, true>(op, stream);
    }
#define P2L_DISPATCH(bn)                                                          \
    case bn:                                                                      \
        return op.mode == EPI_FWD ? launch_t<bn, EPI_FWD, false>(op, stream)      \
                                  : launch_t<bn, EPI_BWD, false>(op, stream);
    switch (op.BN) {
        P2L_DISPATCH(16)
        P2L_DISPATCH(64)
        P2L_DISPATCH(128)
        P2L_DISPATCH(256)
    }
#undef P2L_DISPATCH
    set_error("conv_op_launch: unsupported BN=%d", op.BN);
    return -1;
}

}  // namespace p2l
