// Device-resident inner loop of the inversion: `steps` repetitions of what
// /root/reference pix2latent/optimizer/closure.py:38-71 does per mini-batch —
//     hooks (function_hooks.py:10-27 Clamp)   -> inner_pre_kernel
//     model forward, loss, loss.mean().backward()  -> BigGAN::forward / Lpips::loss_forward / BigGAN::backward
//     optimizer.step (torch.optim.Adam, one param group per latent tensor, variable_manager.py:231-238)
//                                              -> adam_kernel
// with no host round trip in between: the per-step losses and (optionally) the tracked inputs
// (base_optimizer.py:105-106) are recorded in device buffers, the Adam step counter lives in device
// memory, so one step is a STATIC launch sequence that is captured once into a CUDA graph and
// replayed steps-1 times.
#include "optim.h"

#include <cmath>

namespace p2l {

// counters[0] = Adam step count t (persists across calls), counters[1] = iteration within this call
// single block: the counter bump at the end cannot race with a read

// before the forward: record the tracked inputs (the reference clones them BEFORE the hooks run,
// base_optimizer.py:94-97 then closure.py:42-44), then the Clamp hook in place
__global__ void inner_pre_kernel(float* __restrict__ z, float* __restrict__ c, int nz, int nc, float clamp_z, float clamp_c,
                                 float* __restrict__ z_hist, float* __restrict__ c_hist, const int* __restrict__ counters) {
    const int it = counters[1];
    for (int i = threadIdx.x; i < nz + nc; i += blockDim.x) {
        const bool isz = i < nz;
        float* p = isz ? z + i : c + (i - nz);
        float v = *p;
        if (isz) { if (z_hist) z_hist[(long)it * nz + i] = v; }
        else { if (c_hist) c_hist[(long)it * nc + (i - nz)] = v; }
        const float lim = isz ? clamp_z : clamp_c;
        if (lim > 0.f) {
            v = fminf(fmaxf(v, -lim), lim);
            *p = v;
        }
    }
}

// torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False), torch/optim/adam.py
// _single_tensor_adam / _multi_tensor_adam:
//   m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, value=1-b2)
//   step_size = lr / (1 - b1^t); denom = sqrt(v) / sqrt(1 - b2^t) + eps; p.addcdiv_(m, denom, value=-step_size)
// (bias corrections in double on the host there, in double here)
__global__ void adam_kernel(float* __restrict__ z, float* __restrict__ c, const float* __restrict__ dz, const float* __restrict__ dc,
                            float* __restrict__ m, float* __restrict__ v, int nz, int nc, float lr_z, float lr_c, float beta1,
                            float beta2, float eps, int* __restrict__ counters, const float* __restrict__ loss,
                            float* __restrict__ loss_hist, int b) {
    const int t = counters[0] + 1;
    const int it = counters[1];
    const double bc1 = 1.0 - pow((double)beta1, (double)t);
    const double bc2 = 1.0 - pow((double)beta2, (double)t);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float ss_z = (float)((double)lr_z / bc1), ss_c = (float)((double)lr_c / bc1);
    const float w1 = 1.f - beta1, w2 = 1.f - beta2;
    for (int i = threadIdx.x; i < nz + nc; i += blockDim.x) {
        const bool isz = i < nz;
        float* p = isz ? z + i : c + (i - nz);
        const float g = isz ? dz[i] : dc[i - nz];
        float mi = m[i], vi = v[i];
        mi = mi + w1 * (g - mi);
        vi = vi * beta2 + w2 * g * g;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        *p = *p - (isz ? ss_z : ss_c) * (mi / denom);
    }
    if (loss_hist) {
        for (int i = threadIdx.x; i < b; i += blockDim.x) loss_hist[(long)it * b + i] = loss[i];
    }
    __syncthreads();  // every thread has read the counters
    if (threadIdx.x == 0) {
        counters[0] = t;
        counters[1] = it + 1;
    }
}

void k_inner_pre(float* z, float* c, int nz, int nc, float clamp_z, float clamp_c, float* z_hist, float* c_hist,
                 const int* counters, cudaStream_t st) {
    inner_pre_kernel<<<1, 1024, 0, st>>>(z, c, nz, nc, clamp_z, clamp_c, z_hist, c_hist, counters);
    count_launch();
}
void k_adam(float* z, float* c, const float* dz, const float* dc, float* m, float* v, int nz, int nc, const p2l_adam_config& cfg,
            int* counters, const float* loss, float* loss_hist, int b, cudaStream_t st) {
    adam_kernel<<<1, 1024, 0, st>>>(z, c, dz, dc, m, v, nz, nc, cfg.lr_z, cfg.lr_c, cfg.beta1, cfg.beta2, cfg.eps, counters, loss,
                                    loss_hist, b);
    count_launch();
}

InnerLoop::~InnerLoop() {
    if (stream) cudaStreamSynchronize(stream);
    if (exec) cudaGraphExecDestroy(exec);
    if (ev_in) cudaEventDestroy(ev_in);
    if (ev_out) cudaEventDestroy(ev_out);
    if (stream) cudaStreamDestroy(stream);
}

int InnerLoop::ensure_stream() {
    if (stream) return 0;
    P2L_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    P2L_CUDA_CHECK(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
    P2L_CUDA_CHECK(cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming));
    return 0;
}

float* InnerLoop::scratch_for(int b, int zd, int cd) {
    auto it = scratch.find(b);
    if (it != scratch.end()) return it->second;
    float* p = ar.alloc<float>((size_t)b * (1 + zd + cd));
    if (!p) return nullptr;
    scratch[b] = p;
    return p;
}

int biggan_optimize(BigGAN& g, Lpips& l, Target& t, InnerLoop& loop, int b, int steps, float* z, float* c, const float* dloss,
                    float grad_scale, const p2l_adam_config& cfg, float* mv, int* counters, float* loss_hist, float* z_hist,
                    float* c_hist, float* img, int use_graph, cudaStream_t caller) {
    const int zd = g.cfg.z_dim, cd = g.cfg.class_embed_dim;
    const int nz = b * zd, nc = b * cd;
    float* scr = loop.scratch_for(b, zd, cd);
    if (!scr) return -1;
    float *loss = scr, *dz = scr + b, *dc = scr + b + nz;
    float *m = mv, *v = mv + nz + nc;
    loop.graph_used = 0;

    auto one_step = [&](cudaStream_t st) -> int {
        k_inner_pre(z, c, nz, nc, cfg.clamp_z, cfg.clamp_c, z_hist, c_hist, counters, st);
        if (g.forward(b, z, c, nullptr, st)) return -1;
        if (l.loss_forward(t, b, g.last_image(b), loss, 1, st)) return -1;
        float* dimg = l.unit_grad(t, b);
        if (!dimg) return -1;
        if (g.backward(b, dimg, dz, dc, st, grad_scale, dloss)) return -1;
        k_adam(z, c, dz, dc, m, v, nz, nc, cfg, counters, loss, loss_hist, b, st);
        P2L_CUDA_CHECK(cudaGetLastError());
        return 0;
    };

    cudaStream_t st = caller;
    const bool graph = use_graph && steps > 2;
    if (graph) {
        // the legacy default stream cannot be captured: run on an internal stream ordered after / before the caller's
        if (loop.ensure_stream()) return -1;
        P2L_CUDA_CHECK(cudaStreamSynchronize(loop.stream));  // the previous call's graph launches are done: its exec may go
        if (loop.exec) { cudaGraphExecDestroy(loop.exec); loop.exec = nullptr; }
        P2L_CUDA_CHECK(cudaEventRecord(loop.ev_in, caller));
        P2L_CUDA_CHECK(cudaStreamWaitEvent(loop.stream, loop.ev_in, 0));
        st = loop.stream;
    }
    // iteration counter of this call
    P2L_CUDA_CHECK(cudaMemsetAsync(counters + 1, 0, sizeof(int), st));
    int done = 0;
    // step 0 eagerly: builds the per-batch plans, sets the kernels' shared-memory attributes
    if (steps > 0) {
        if (one_step(st)) return -1;
        done = 1;
    }
    if (graph) {
        const long n0 = launch_count();
        cudaGraph_t gr = nullptr;
        bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess;
        if (ok) {
            const int rc = one_step(st);
            const cudaError_t e = cudaStreamEndCapture(st, &gr);
            ok = (rc == 0) && (e == cudaSuccess) && gr != nullptr;
        }
        const long per_step = launch_count() - n0;
        if (ok) ok = cudaGraphInstantiate(&loop.exec, gr, 0) == cudaSuccess;
        if (gr) cudaGraphDestroy(gr);
        if (ok) {
            add_launches(-per_step);  // the captured pass did not execute
            for (; done < steps; ++done) {
                if (cudaGraphLaunch(loop.exec, st) != cudaSuccess) { ok = false; break; }
                add_launches(per_step);
            }
            loop.graph_used = ok ? 1 : 0;
        }
        if (!ok) {
            // capture / instantiate / launch refused: clear the sticky-less error and finish eagerly
            cudaGetLastError();
            if (loop.exec) { cudaGraphExecDestroy(loop.exec); loop.exec = nullptr; }
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            cudaStreamIsCapturing(st, &cs);
            if (cs != cudaStreamCaptureStatusNone) { cudaGraph_t junk = nullptr; cudaStreamEndCapture(st, &junk); if (junk) cudaGraphDestroy(junk); cudaGetLastError(); }
        }
    }
    for (; done < steps; ++done)
        if (one_step(st)) return -1;
    if (img && steps > 0) {
        const int R = g.H_out;
        P2L_CUDA_CHECK(cudaMemcpyAsync(img, g.last_image(b), (size_t)b * 3 * R * R * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    if (graph) {
        P2L_CUDA_CHECK(cudaEventRecord(loop.ev_out, st));
        P2L_CUDA_CHECK(cudaStreamWaitEvent(caller, loop.ev_out, 0));
    }
    return 0;
}

}  // namespace p2l
