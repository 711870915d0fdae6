// Native LPIPS + pixel loss (see lpips.cu).
#pragma once
#include <memory>

#include "model_common.h"
#include "p2l.h"

namespace p2l {

struct LpipsPlan;
struct Lpips;

struct Target {
    Lpips* m = nullptr;
    int H = 0, W = 0, rec_type = 1;
    float rec_w = 1.f, per_w = 10.f, sumW = 0.f;
    Arena ar;
    float *target = nullptr, *weight = nullptr, *mask = nullptr, *wsum = nullptr, *total = nullptr;
    float* tfeat[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float* wadj[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int fh[5] = {0, 0, 0, 0, 0}, fw[5] = {0, 0, 0, 0, 0};
};

struct Lpips {
    Ctx* ctx = nullptr;
    int net = 0;
    TensorStage stage;
    Arena weights;
    bool finalized = false;
    struct LConv {
        int Cin, Cout, k, stride, pad;
        int pool_before, pool_k, pool_s;
        int feat;  // feature index if this conv's relu output is an LPIPS layer, else -1
        int Kp;    // padded K of the first layer
        std::string name;
        act_t *w, *wt;
        float* bias;
    };
    std::vector<LConv> convs;
    int nfeat = 0;
    float* lin[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int chns[5] = {0, 0, 0, 0, 0};
    std::map<long, std::shared_ptr<LpipsPlan>> plans;

    int finalize();
    int feature_dims(int H, int W, int* fh, int* fw) const;
    LpipsPlan* plan(int b, int H, int W);
    int features(LpipsPlan& P, const float* img, cudaStream_t st);
    Target* make_target(const float* target, const float* weight, const float* mask, int H, int W, int rec_type,
                        float rec_weight, float per_weight, cudaStream_t st);
    int loss_forward(Target& T, int b, const float* img, float* loss, int want_grad, cudaStream_t st);
    int loss_forward_multi(Target* const* Ts, int b, const float* img, float* loss, int want_grad, cudaStream_t st);
    int loss_backward(Target& T, int b, const float* dloss, float* dimg, cudaStream_t st);
    float* unit_grad(Target& T, int b);
    double flops(int b, int H, int W, int backward);
    ~Lpips();
};

}  // namespace p2l
