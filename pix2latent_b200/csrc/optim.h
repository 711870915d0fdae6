// Device-resident inner loop: Clamp hook + fused step + Adam, optionally replayed as a CUDA graph (optim.cu).
#pragma once
#include <map>

#include "biggan.h"
#include "lpips.h"

namespace p2l {

// per-generator resources of the inner loop: capture stream, the instantiated graph of the last call,
// per-batch scratch (loss[b], dz[b,zd], dc[b,cd])
struct InnerLoop {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    cudaGraphExec_t exec = nullptr;
    Arena ar;
    std::map<int, float*> scratch;
    int graph_used = 0;  // 1 when the last call ran through a CUDA graph
    int ensure_stream();
    float* scratch_for(int b, int zd, int cd);
    ~InnerLoop();
};

void k_inner_pre(float* z, float* c, int nz, int nc, float clamp_z, float clamp_c, float* z_hist, float* c_hist,
                 const int* counters, cudaStream_t st);
// one Adam update of (z, c) with gradients (dz, dc); m, v: [nz + nc] each; counters[0] = step count (bumped),
// counters[1] = row of loss_hist[., b] that receives loss[b] (bumped)
void k_adam(float* z, float* c, const float* dz, const float* dc, float* m, float* v, int nz, int nc, const p2l_adam_config& cfg,
            int* counters, const float* loss, float* loss_hist, int b, cudaStream_t st);

int biggan_optimize(BigGAN& g, Lpips& l, Target& t, InnerLoop& loop, int b, int steps, float* z, float* c, const float* dloss,
                    float grad_scale, const p2l_adam_config& cfg, float* mv, int* counters, float* loss_hist, float* z_hist,
                    float* c_hist, float* img, int use_graph, cudaStream_t caller);

}  // namespace p2l
