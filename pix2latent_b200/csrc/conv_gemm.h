// Host-side description of one tcgen05 implicit-GEMM launch (see conv_gemm.cuh).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "conv_gemm.cuh"

namespace p2l {

struct ConvDesc {
    // A: bf16 NHWC tensor read by TMA
    const void* A = nullptr;
    int A_N = 0, A_H = 0, A_W = 0, A_C = 0;
    int a_c0 = 0;  // first channel used
    int Cin = 0;   // channels used per tap (multiple of 64)
    // B: bf16 [batch][Cout][kh*kw*Cin]
    const void* B = nullptr;
    int Cout = 0;
    int B_batch = 0;  // 0: shared weights; >0: one matrix per image
    int kh = 1, kw = 1, pad_h = 0, pad_w = 0;
    // output pixel grid
    int NI = 0, H = 0, W = 0;
    int BN = 128;
    int mode = EPI_FWD;
    ConvGemmParams epi{};  // only the epilogue fields are read from here
    int sg = 0;            // StyleGAN2 modulated-convolution epilogue (FLAVOR_SG kernels, sg_epilogue.cuh)
};

struct ConvOp {
    CUtensorMap tmA, tmB;
    OutMaps tmO;
    int tma_out;  // epilogue outputs through TMA bulk stores
    ConvGemmParams p;
    int BN, mode, grid;
    int deep;  // one CTA per SM, full-depth pipeline (few tiles, long K)
    int flavor;   // FLAVOR_PLAIN | FLAVOR_ROWFUSE (attention: row-wise softmax fusions) | FLAVOR_SG (StyleGAN2 epilogues)
    int halo;  // 0: per-tap A loads; 10 / 16: halo-patch kernel with that patch row pitch
    int halo_smem;  // dynamic shared memory of the halo kernel for this plan
    int stat_parts; // partial slots per image of the BN-gradient sums this launch fills
    double flops;  // algorithmic 2*M*N*K of this launch
};

// returns 0 on success; on failure sets the thread-local error string
int conv_op_build(ConvOp* op, const ConvDesc& d);
int conv_op_launch(const ConvOp& op, cudaStream_t stream);
// upper bound of ConvOp::stat_parts for an H x W output grid (either tiling: 16x8 or the halo kernel's 8x16)
int conv_stat_parts_max(int H, int W);

void set_error(const char* fmt, ...);
const char* get_error();
int num_sms();
void count_launch();  // every kernel launch of the library is counted (bench.py gpu_launches)
long launch_count();
void add_launches(long n);  // graph replays: the captured sequence's launches per replay
// tuning / experiment switches: "halo" (patch pitch 10 | 16), "halo_mode" (0 off, 1 resident-weight layers, 2 all 3x3), "halo_bo" (0/1), "tma_out", "tma_kmax", "grad_scale"
// Static loss scale carried by 16-bit gradients (act_type.h); default kGradScale, option "grad_scale".
float grad_scale();
void set_option(const char* key, int value);
int get_option(const char* key);
void profile_enable(int on);
int profile_read(double* conv_ms, long* conv_launches, double* conv_flops);
int profile_get(int i, float* ms, double* flops, int* info);

#define P2L_CUDA_CHECK(expr)                                                             \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::p2l::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                             __FILE__, __LINE__);                                        \
            return -1;                                                                   \
        }                                                                                \
    } while (0)

}  // namespace p2l
