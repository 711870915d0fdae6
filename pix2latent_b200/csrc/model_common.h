// Shared host-side helpers for the native models: device arena, named-tensor staging.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "conv_gemm.h"
#include "kernels.h"

namespace p2l {

struct Ctx {
    int device = 0;
    int sm_count = 0;
};

// Simple owning list of cudaMalloc allocations.
struct Arena {
    std::vector<void*> ptrs;
    size_t total = 0;
    bool failed = false;
    template <typename T>
    T* alloc(size_t n, bool zero = false) {
        void* p = nullptr;
        const size_t bytes = ((n * sizeof(T) + 255) / 256) * 256 + 256;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            failed = true;
            set_error("cudaMalloc of %zu bytes failed", bytes);
            return nullptr;
        }
        if (zero) cudaMemset(p, 0, bytes);
        ptrs.push_back(p);
        total += bytes;
        return static_cast<T*>(p);
    }
    void release() {
        for (void* p : ptrs) cudaFree(p);
        ptrs.clear();
        total = 0;
    }
    ~Arena() { release(); }
};

inline act_t host_f2bf(float f) { return f2a(f); }

template <typename T>
inline T* upload(Arena& ar, const std::vector<T>& h) {
    T* d = ar.alloc<T>(h.size());
    if (d) cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

struct TensorStage {
    std::map<std::string, std::vector<float>> t;
    int set(const char* name, const float* data, long numel) {
        std::vector<float>& v = t[name];
        v.resize(numel);
        if (cudaMemcpy(v.data(), data, numel * sizeof(float), cudaMemcpyDefault) != cudaSuccess) {
            set_error("set_tensor(%s): copy of %ld floats failed", name, numel);
            return -1;
        }
        return 0;
    }
    const std::vector<float>* get(const std::string& name, long expect) {
        auto it = t.find(name);
        if (it == t.end()) {
            set_error("missing tensor '%s'", name.c_str());
            return nullptr;
        }
        if (expect >= 0 && (long)it->second.size() != expect) {
            set_error("tensor '%s' has %zu elements, expected %ld", name.c_str(), it->second.size(), expect);
            return nullptr;
        }
        return &it->second;
    }
};

// conv weight [Cout, Cin, kh, kw] fp32 -> forward GEMM operand [Cout][(r*kw+s)*Cin + c] bf16
inline std::vector<act_t> pack_conv_fwd(const std::vector<float>& w, int Cout, int Cin, int kh, int kw,
                                                int Cout_keep = -1) {
    if (Cout_keep < 0) Cout_keep = Cout;
    std::vector<act_t> o((size_t)Cout_keep * kh * kw * Cin);
    for (int oc = 0; oc < Cout_keep; ++oc)
        for (int r = 0; r < kh; ++r)
            for (int s = 0; s < kw; ++s)
                for (int c = 0; c < Cin; ++c)
                    o[((size_t)oc * kh * kw + r * kw + s) * Cin + c] =
                        host_f2bf(w[(((size_t)oc * Cin + c) * kh + r) * kw + s]);
    return o;
}
// dgrad operand [Cin][(r'*kw+s')*Cout + o] = W[o, c, kh-1-r', kw-1-s'] bf16
inline std::vector<act_t> pack_conv_dgrad(const std::vector<float>& w, int Cout, int Cin, int kh, int kw) {
    std::vector<act_t> o((size_t)Cin * kh * kw * Cout);
    for (int c = 0; c < Cin; ++c)
        for (int r = 0; r < kh; ++r)
            for (int s = 0; s < kw; ++s)
                for (int oc = 0; oc < Cout; ++oc)
                    o[((size_t)c * kh * kw + r * kw + s) * Cout + oc] =
                        host_f2bf(w[(((size_t)oc * Cin + c) * kh + (kh - 1 - r)) * kw + (kw - 1 - s)]);
    return o;
}

inline int pick_bn(int Cout) {
    if (Cout <= 16) return 16;
    if (Cout % 256 == 0) return 256;
    if (Cout % 128 == 0) return 128;
    return 64;
}

}  // namespace p2l
