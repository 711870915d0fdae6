// Native BigGAN-deep generator: forward + dgrad-only backward to the latent (z, c).
// Restates pytorch_pretrained_biggan's Generator as reached through
// /root/reference pix2latent/model/biggan.py:50-58 (see oracle/biggan.py for the CPU statement).
#pragma once
#include <memory>

#include "model_common.h"
#include "p2l.h"

namespace p2l {

struct BigGANPlan;

struct BigGAN {
    Ctx* ctx = nullptr;
    p2l_biggan_config cfg{};
    TensorStage stage;
    Arena weights;
    bool finalized = false;

    struct BN { int C, off; bool cond; };
    std::vector<BN> bns;  // conditional ones first (concatenated tables), the final unconditional last
    int C_cond = 0, C_all = 0, cdim = 0;
    float *WT_as = nullptr, *bias_as = nullptr;  // cond -> BN (gain | offset) GEMV with the statistics folded in: [cdim][2*C_cond], [2*C_cond]
    float *mean = nullptr, *inv_std = nullptr;  // [C_all]
    float *unc_weight = nullptr, *unc_bias = nullptr;
    float *Wcat = nullptr;  // [2*C_cond, cdim] rows: Ws then Wo (for dcond)
    float *genz_W = nullptr, *genz_WT = nullptr, *genz_b = nullptr;
    int genz_J = 0, C0 = 0;

    struct Block {
        int in, out, mid, Hin, Hout;
        bool up;
        int bn[4];
        act_t *w[4], *wt[4];
        float* bias[4];
    };
    std::vector<Block> blocks;
    struct Attn {
        int C = 0, H = 0, dq = 0, dv = 0;
        act_t *wqkv = nullptr, *wqkv_t = nullptr, *wo = nullptr, *wo_t = nullptr;
        float* gamma = nullptr;
    } attn;
    int final_bn = -1;
    int C_last = 0, H_out = 0;
    act_t *wrgb = nullptr, *wrgb_t = nullptr;  // wrgb: tap-expanded head [27 = tap*3 + o][C_last]
    float* brgb = nullptr;

    std::map<int, std::shared_ptr<BigGANPlan>> plans;
    BigGANPlan* last_plan = nullptr;

    int finalize();
    BigGANPlan* plan(int b);
    int forward(int b, const float* z, const float* c, float* img, cudaStream_t st);
    int backward(int b, const float* dimg, float* dz, float* dc, cudaStream_t st, float scale = 1.f,
                 const float* row_scale = nullptr);
    const float* last_image(int b);
    size_t device_bytes();
    double flops(int b, int backward);
    ~BigGAN();
};

}  // namespace p2l
