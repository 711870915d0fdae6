// Kernel-level debug entry points (include/p2l_debug.h).
#include "p2l_debug.h"

#include "conv_gemm.h"

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
    e.stat0 = a->stat0; e.stat1 = a->stat1; e.stat_stride = a->stat_stride;
    e.addin = static_cast<const act_t*>(a->addin); e.addin_C = a->addin_C;
    e.addin_climit = a->addin_climit; e.addin_pool = a->addin_pool;
    e.dx = static_cast<act_t*>(a->dx); e.dx_C = a->dx_C;
    e.dx_f32 = a->dx_f32; e.dx_f32_C = a->dx_f32_C;
    e.rowstat = a->rowstat; e.rowstat_in = a->rowstat_in; e.rowstat_nt = a->rowstat_nt;
    e.rowsub = a->rowsub; e.mulin = static_cast<const act_t*>(a->mulin); e.mulin_C = a->mulin_C;
    d.splitk_ws = a->splitk_ws; d.splitk_ws_floats = a->splitk_ws_floats;
    ConvOp op;
    if (conv_op_build(&op, d)) return -1;
    return conv_op_launch(op, static_cast<cudaStream_t>(cuda_stream));
}

P2L_EXPORT const char* p2l_last_error(void) { return get_error(); }
P2L_EXPORT void p2l_debug_set_option(const char* key, int value) { set_option(key, value); }
P2L_EXPORT int p2l_debug_get_option(const char* key) { return get_option(key); }

P2L_EXPORT int p2l_debug_profile_get(int i, float* ms, double* flops, int* info) { return profile_get(i, ms, flops, info); }
