/* Kernel-level C-ABI used by the parity tests (tests/test_conv_gemm_gpu.py): one launch of the
 * tcgen05 implicit-GEMM convolution with an explicit epilogue description. Not part of the
 * drop-in boundary (that is include/p2l.h); exported so that every kernel can be checked against
 * the oracle in isolation. All pointers are device pointers owned by the caller. */
#ifndef P2L_DEBUG_H
#define P2L_DEBUG_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct p2l_conv_args {
    /* A: 16-bit (p2l_act_dtype) NHWC */
    const void* A;
    int A_N, A_H, A_W, A_C, a_c0, Cin;
    /* B: 16-bit [batch][Cout][kh*kw*Cin] */
    const void* B;
    int Cout, B_batch, kh, kw, pad_h, pad_w;
    /* output pixel grid, tile N, mode (0 fwd, 1 bwd) */
    int NI, H, W, BN, mode;
    /* forward epilogue */
    float alpha;
    const float* alpha_ptr;
    const float* bias;
    const void* resid;
    int resid_C, resid_shift;
    void* raw;
    int raw_C;
    float* raw_f32;
    int raw_f32_C;
    const float* aff_a;
    const float* aff_s;
    int aff_stride, relu;
    void* act;
    int act_C, act_up;
    void* act_lo;
    float* img_nchw;
    /* backward epilogue */
    const void* saved;
    int saved_C;
    float* stat0; /* [NI, stat_stride]: sum_pix dpre, sum_pix dpre*saved (overwritten; summed in a fixed order) */
    float* stat1;
    int stat_stride;
    const void* addin;
    int addin_C, addin_climit, addin_pool;
    void* dx;
    int dx_C;
    float* dx_f32;
    int dx_f32_C;
    /* forward epilogue, row-wise softmax fusions (attention) */
    float* rowstat;
    const float* rowstat_in;
    int rowstat_nt;
    const float* rowsub;
    const void* mulin;
    int mulin_C;
    /* transposed 16-bit copy of the main output (fwd: raw, bwd: dx), channels [outT_c0, outT_c1): [NI][c][H*W] */
    void* outT;
    int outT_c0, outT_c1;
    /* walk the tiles from the last to the first (same result) */
    int tile_reverse;
} p2l_conv_args;

/* returns 0 on success, <0 on error (see p2l_last_error) */
int p2l_debug_conv(const p2l_conv_args* args, void* cuda_stream);

/* experiment switches of the kernel library: "halo" (0 = per-tap A loads, 10 / 16 = halo-patch
 * kernel with that patch row pitch), "halo_bo" (descriptor base-offset mode) */
void p2l_debug_set_option(const char* key, int value);
int p2l_debug_get_option(const char* key);
/* i-th tensor-core launch recorded since p2l_profile_enable(1): duration (ms), algorithmic FLOPs,
 * info[7] = {BN, mode, halo, grid, M, N, K}; call before p2l_profile_read (which resets) */
int p2l_debug_profile_get(int i, float* ms, double* flops, int* info);

#ifdef __cplusplus
}
#endif
#endif
