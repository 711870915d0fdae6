// Kernel-level debug entry points (include/p2l_debug.h).
#include "p2l_debug.h"

#include "conv_gemm.h"
#include "kernels.h"

#include <vector>

#define P2L_EXPORT extern "C" __attribute__((visibility("default")))

using namespace p2l;

P2L_EXPORT int p2l_debug_conv(const p2l_conv_args* a, void* cuda_stream) {
    ConvDesc d;
    d.A = a->A; d.A_N = a->A_N; d.A_H = a->A_H; d.A_W = a->A_W; d.A_C = a->A_C;
    d.a_c0 = a->a_c0; d.Cin = a->Cin;
    d.B = a->B; d.Cout = a->Cout; d.B_batch = a->B_batch;
    d.kh = a->kh; d.kw = a->kw; d.pad_h = a->pad_h; d.pad_w = a->pad_w;
    d.NI = a->NI; d.H = a->H; d.W = a->W; d.BN = a->BN; d.mode = a->mode;
    ConvGemmParams& e = d.epi;
    e.alpha = a->alpha; e.alpha_ptr = a->alpha_ptr; e.bias = a->bias;
    e.resid = static_cast<const act_t*>(a->resid); e.resid_C = a->resid_C; e.resid_shift = a->resid_shift;
    e.raw = static_cast<act_t*>(a->raw); e.raw_C = a->raw_C;
    e.raw_f32 = a->raw_f32; e.raw_f32_C = a->raw_f32_C;
    e.aff_a = a->aff_a; e.aff_s = a->aff_s; e.aff_stride = a->aff_stride; e.relu = a->relu;
    e.act = static_cast<act_t*>(a->act); e.act_C = a->act_C; e.act_up = a->act_up;
    e.act_lo = static_cast<act_t*>(a->act_lo); e.img_nchw = a->img_nchw;
    e.saved = static_cast<const act_t*>(a->saved); e.saved_C = a->saved_C;
    // BN-gradient sums: the kernel fills per-tile partial slots (ConvGemmParams::statp); this entry point keeps the
    // (stat0, stat1, stat_stride) view of the result by running the fixed-order reduction itself
    float* statp = nullptr;
    StatSeg* dsegs = nullptr;
    const int parts = conv_stat_parts_max(a->H, a->W);
    if (a->stat0 && a->stat1) {
        if (cudaMalloc(&statp, (size_t)a->NI * parts * 2 * a->Cout * sizeof(float)) != cudaSuccess) { set_error("debug_conv: cudaMalloc failed"); return -1; }
        e.statp = statp; e.statp_parts = parts; e.statp_C = a->Cout;
    }
    e.outT = static_cast<act_t*>(a->outT); e.outT_c0 = a->outT_c0; e.outT_c1 = a->outT_c1;
    e.addin = static_cast<const act_t*>(a->addin); e.addin_C = a->addin_C;
    e.addin_climit = a->addin_climit; e.addin_pool = a->addin_pool;
    e.dx = static_cast<act_t*>(a->dx); e.dx_C = a->dx_C;
    e.dx_f32 = a->dx_f32; e.dx_f32_C = a->dx_f32_C;
    e.rowstat = a->rowstat; e.rowstat_in = a->rowstat_in; e.rowstat_nt = a->rowstat_nt;
    e.rowsub = a->rowsub; e.mulin = static_cast<const act_t*>(a->mulin); e.mulin_C = a->mulin_C;
    e.tile_reverse = a->tile_reverse;
    ConvOp op;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    int rc = conv_op_build(&op, d);
    if (rc == 0) rc = conv_op_launch(op, st);
    if (rc == 0 && statp) {
        if (a->Cout % 32) { set_error("debug_conv: statistics need Cout %% 32 == 0"); rc = -1; }
        else {
            std::vector<StatSeg> segs;
            for (int c0 = 0; c0 < a->Cout; c0 += 32) segs.push_back({statp, op.stat_parts, parts, a->Cout, 0, c0, 0});
            if (cudaMalloc(&dsegs, segs.size() * sizeof(StatSeg)) != cudaSuccess) { set_error("debug_conv: cudaMalloc failed"); rc = -1; }
            else {
                cudaMemcpyAsync(dsegs, segs.data(), segs.size() * sizeof(StatSeg), cudaMemcpyHostToDevice, st);
                k_stat_reduce(dsegs, (int)segs.size(), a->stat0, a->stat1, a->stat_stride, a->NI, st);
            }
        }
    }
    if (statp || dsegs) {
        cudaStreamSynchronize(st);
        cudaFree(statp);
        cudaFree(dsegs);
    }
    return rc;
}

P2L_EXPORT const char* p2l_last_error(void) { return get_error(); }
P2L_EXPORT void p2l_debug_set_option(const char* key, int value) { set_option(key, value); }
P2L_EXPORT int p2l_debug_get_option(const char* key) { return get_option(key); }

P2L_EXPORT int p2l_debug_profile_get(int i, float* ms, double* flops, int* info) { return profile_get(i, ms, flops, info); }
