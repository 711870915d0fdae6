// Native StyleGAN2 generator (rosinality Generator(size, 512, 8, channel_multiplier=2), z search):
// forward + backward to z. Restates what /root/reference pix2latent/model/stylegan2.py:116-119
// runs; CPU statement in oracle/stylegan2.py.
#pragma once
#include <memory>

#include "model_common.h"
#include "p2l.h"
#include "sg2_kernels.h"

namespace p2l {

struct SG2Plan;

struct SG2 {
    Ctx* ctx = nullptr;
    p2l_sg2_config cfg{};
    TensorStage stage;
    Arena weights;
    bool finalized = false;
    int sdim = 512, n_mlp = 8, log_size = 0, num_layers = 0;  // num_layers = StyledConv count = noise layers

    // mapping network
    std::vector<float*> map_WT, map_W, map_b;  // [in][out], [out][in], bias * lr_mul
    float map_scale = 0.f;
    // all modulation affines concatenated: S = sum of Cin over StyledConvs then ToRGBs
    int S = 0;
    float *affT = nullptr, *aff = nullptr, *aff_b = nullptr;  // [512][S], [S][512], [S]
    struct Conv {   // StyledConv
        int Cin, Cout, Hin, Hout, up;
        int s_off, dm_off;
        act_t *w, *wt;      // shared weights (scale folded in), forward / dgrad GEMM operands
        float *wsqT, *wsq;          // [Cin][Cout], [Cout][Cin]: sum_k (scale W)^2
        float *noise_w, *bias;
    };
    struct Rgb { int Cin, H, s_off; float* Wr; float* bias; float scale; };
    std::vector<Conv> convs;
    std::vector<Rgb> rgbs;
    int DM = 0;  // sum of Cout over StyledConvs
    act_t* const_in = nullptr;  // [4,4,C0] NHWC
    std::map<int, std::shared_ptr<SG2Plan>> plans;

    int finalize();
    SG2Plan* plan(int b);
    int forward(int b, const float* z, const float* const* noise, float* img, cudaStream_t st);
    int backward(int b, const float* dimg, float* dz, cudaStream_t st, float scale = 1.f, const float* row_scale = nullptr);
    // w / w+ search (pix2latent/model/stylegan2.py:122-125 forward_w: Generator([w], input_is_latent=True, noise=noises)):
    // latent [b, n_latent, sdim] (layer l reads row l, ToRGB t reads row 2t+1); backward to the latent rows and,
    // when dnoise != null, to every layer's noise image
    int n_latent() const { return 2 * log_size - 2; }
    int style(int b, const float* z, float* w, cudaStream_t st);  // mapping network alone: w = style(z)
    int run_mapping(int b, const float* z, float* zbuf, float* const* h, cudaStream_t st);
    Arena style_ar;
    std::map<int, float*> style_scratch;
    int forward_w(int b, const float* latent, const float* const* noise, float* img, cudaStream_t st);
    int backward_w(int b, const float* dimg, float* dlatent, float* const* dnoise, cudaStream_t st, float scale = 1.f,
                   const float* row_scale = nullptr);
    int synth(SG2Plan& P, int b, const float* const* noise, float* img, cudaStream_t st);
    int synth_bwd(SG2Plan& P, int b, const float* dimg, float* const* dnoise, float out_scale, const float* row_scale,
                  cudaStream_t st);
    const float* last_image(int b);
    ~SG2();
};

}  // namespace p2l
