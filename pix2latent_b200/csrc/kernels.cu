// Non-contraction kernels of the inversion step (see kernels.h). These are the HBM-bound glue
// between the tcgen05 convolutions: latent-side GEMVs, pooling, softmax, the LPIPS distance and
// the loss reductions (warp-shuffle + one atomic per block).
#include "kernels.h"
#include "conv_gemm.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace p2l {

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float b2f(bf16 x) { return a2f(x); }
__device__ __forceinline__ bf16 f2b(float x) { return f2a(x); }

// ============================================================================= latent side
__global__ void concat_cond_kernel(const float* z, const float* c, float* cond, int b, int zd, int cd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int D = zd + cd;
    if (i >= b * D) return;
    const int bi = i / D, k = i % D;
    cond[i] = k < zd ? z[bi * zd + k] : c[bi * cd + (k - zd)];
}
void k_concat_cond(const float* z, const float* c, float* cond, int b, int zd, int cd, cudaStream_t st) {
    concat_cond_kernel<<<cdiv((long)b * (zd + cd), 256), 256, 0, st>>>(z, c, cond, b, zd, cd); count_launch();
}

// thread per channel; weights stored transposed ([cdim][C]) so that a warp reads 128 contiguous
// bytes per k; cond staged in shared memory (broadcast reads); 8 samples per register pass
// ---- small-batch GEMV family: y[bi][j] = bias[j] + sum_k x[bi][k] * WT[k][j]
// (cond -> BN gains/offsets, cond -> gen_z). Batches of <= 24 rows ride along a single pass over
// the fp32 weights: one thread owns 4 consecutive outputs j (one 16-byte weight load per k) for a
// quarter of the k range and 24 x 4 accumulators; x is staged transposed and zero-padded in shared
// memory so 24 rows of one k come back in six broadcast LDS.128 — 96 FMAs per 7 memory
// instructions. The four k-quarters sit in one warp and meet through two xor-shuffles.
constexpr int kGvB = 24;
constexpr int kGvU = 8;   // weight loads in flight per thread

__device__ __forceinline__ void gv_stage_x(const float* __restrict__ x, int ldx, int b0, int b, int K, int k0, float* xs) {
    // xs[k][24] = x[b0 + i][k0 + k], zero beyond row b
    for (int i = threadIdx.x; i < K * kGvB; i += blockDim.x) {
        const int k = i / kGvB, bi = i - k * kGvB;
        xs[i] = (b0 + bi < b) ? x[(long)(b0 + bi) * ldx + k0 + k] : 0.f;
    }
}
__device__ __forceinline__ void gv_fma(float (&acc)[kGvB][4], const float4* xs4, int k, const float4 w) {
#pragma unroll
    for (int q = 0; q < kGvB / 4; ++q) {
        const float4 xv = xs4[k * (kGvB / 4) + q];
        const float xr[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            acc[q * 4 + r][0] = fmaf(xr[r], w.x, acc[q * 4 + r][0]);
            acc[q * 4 + r][1] = fmaf(xr[r], w.y, acc[q * 4 + r][1]);
            acc[q * 4 + r][2] = fmaf(xr[r], w.z, acc[q * 4 + r][2]);
            acc[q * 4 + r][3] = fmaf(xr[r], w.w, acc[q * 4 + r][3]);
        }
    }
}

struct GvEpi {
    // EPI 0: fp32 outputs, columns [0, split) -> out0, [split, J) -> out1, row stride `stride`
    float *out0, *out1;
    int split, stride;
    // EPI 1 (gen_z): 16-bit raw = y and act = relu(a[bi][j % C] * y + s[bi][j % C]), row length J
    const float *a, *s;
    int aff_stride, C;
    bf16 *raw, *act;
};

template <int EPI>
__global__ void __launch_bounds__(128) gemv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ WT,
                                                       const float* __restrict__ bias, int b, int cdim, int J,
                                                       const GvEpi e) {
    extern __shared__ float4 gv_smem[];
    float* xs = reinterpret_cast<float*>(gv_smem);
    const int b0 = blockIdx.y * kGvB;
    gv_stage_x(x, cdim, b0, b, cdim, 0, xs);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kq = lane >> 3;
    const int j = (blockIdx.x * 4 + warp) * 32 + (lane & 7) * 4;
    const bool live = j < J;  // J % 4 == 0
    float acc[kGvB][4];
#pragma unroll
    for (int i = 0; i < kGvB; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    if (live) {
        const float* wp = WT + j;
        for (int k = kq; k < cdim; k += 4 * kGvU) {
            float4 w[kGvU];
#pragma unroll
            for (int u = 0; u < kGvU; ++u)
                w[u] = (k + 4 * u < cdim) ? __ldg(reinterpret_cast<const float4*>(wp + (long)(k + 4 * u) * J)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < kGvU; ++u)
                if (k + 4 * u < cdim) gv_fma(acc, gv_smem, k + 4 * u, w[u]);
        }
    }
#pragma unroll
    for (int i = 0; i < kGvB; ++i) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            acc[i][c] += __shfl_xor_sync(0xffffffffu, acc[i][c], 8);
            acc[i][c] += __shfl_xor_sync(0xffffffffu, acc[i][c], 16);
        }
    }
    if (!live) return;
    const float4 bj = bias ? __ldg(reinterpret_cast<const float4*>(bias + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kGvB; ++i) {
        if ((i & 3) != kq || b0 + i >= b) continue;   // the four k-quarter lanes share the rows
        const int bi = b0 + i;
        const float4 y = make_float4(acc[i][0] + bj.x, acc[i][1] + bj.y, acc[i][2] + bj.z, acc[i][3] + bj.w);
        if constexpr (EPI == 0) {
            float* dst = (j < e.split) ? e.out0 + (long)bi * e.stride + j : e.out1 + (long)bi * e.stride + (j - e.split);
            *reinterpret_cast<float4*>(dst) = y;
        } else {
            const int ch = j % e.C;
            const float4 av = __ldg(reinterpret_cast<const float4*>(e.a + (long)bi * e.aff_stride + ch));
            const float4 sv = __ldg(reinterpret_cast<const float4*>(e.s + (long)bi * e.aff_stride + ch));
            *reinterpret_cast<uint2*>(e.raw + (long)bi * J + j) = make_uint2(pack_act(y.x, y.y), pack_act(y.z, y.w));
            *reinterpret_cast<uint2*>(e.act + (long)bi * J + j) =
                make_uint2(pack_act(fmaxf(fmaf(av.x, y.x, sv.x), 0.f), fmaxf(fmaf(av.y, y.y, sv.y), 0.f)),
                           pack_act(fmaxf(fmaf(av.z, y.z, sv.z), 0.f), fmaxf(fmaf(av.w, y.w, sv.w), 0.f)));
        }
    }
}
// BN affine tables: [a | s][bi][ch] = bias_as[ch] + cond[bi] . WT_as[:, ch]; the BN statistics are folded
// into WT_as / bias_as when the weights are packed (biggan.cu finalize)
void k_cond_affine(const float* cond, const float* WT_as, const float* bias_as, float* a, float* s, int b, int cdim,
                   int C_cond, int stride, cudaStream_t st) {
    GvEpi e{};
    e.out0 = a; e.out1 = s; e.split = C_cond; e.stride = stride;
    const dim3 grid(cdiv(2 * C_cond, 128), cdiv(b, kGvB));
    gemv_fwd_kernel<0><<<grid, 128, (size_t)cdim * kGvB * sizeof(float), st>>>(cond, WT_as, bias_as, b, cdim, 2 * C_cond, e);
    count_launch();
}

__global__ void uncond_affine_kernel(const float* weight, const float* bias, const float* mean,
                                     const float* inv_std, float* a, float* s, int b, int C_cond,
                                     int C_unc, int stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b * C_unc) return;
    const int bi = i / C_unc, ch = i % C_unc;
    const float av = weight[ch] * inv_std[C_cond + ch];
    a[(long)bi * stride + C_cond + ch] = av;
    s[(long)bi * stride + C_cond + ch] = bias[ch] - mean[C_cond + ch] * av;
}
void k_uncond_affine(const float* weight, const float* bias, const float* mean, const float* inv_std,
                     float* a, float* s, int b, int C_cond, int C_unc, int stride, cudaStream_t st) {
    uncond_affine_kernel<<<cdiv((long)b * C_unc, 256), 256, 0, st>>>(weight, bias, mean, inv_std, a, s, b,
                                                                     C_cond, C_unc, stride); count_launch();
}

void k_gen_z(const float* cond, const float* WT, const float* bias, const float* a, const float* s,
             int aff_stride, bf16* raw, bf16* act, int b, int cdim, int J, int C, cudaStream_t st) {
    GvEpi e{};
    e.a = a; e.s = s; e.aff_stride = aff_stride; e.C = C; e.raw = raw; e.act = act;
    const dim3 grid(cdiv(J, 128), cdiv(b, kGvB));
    gemv_fwd_kernel<1><<<grid, 128, (size_t)cdim * kGvB * sizeof(float), st>>>(cond, WT, bias, b, cdim, J, e);
    count_launch();
}

__global__ void bn_grad_finalize_kernel(const float* S0, const float* S1, const float* a, const float* s,
                                        const float* mean, const float* inv_std, float* G, int b, int C,
                                        int stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b * C) return;
    const int bi = i / C, ch = i % C;
    const long o = (long)bi * stride + ch;
    const float av = a[o], sv = s[o];
    const float ds = S0[o];
    const float da = fabsf(av) > 1e-20f ? (S1[o] - sv * ds) / av : 0.f;
    G[(long)bi * 2 * C + ch] = inv_std[ch] * (da - mean[ch] * ds);
    G[(long)bi * 2 * C + C + ch] = ds;
}
void k_bn_grad_finalize(const float* S0, const float* S1, const float* a, const float* s,
                        const float* mean, const float* inv_std, float* G, int b, int C_cond,
                        int stride, cudaStream_t st) {
    bn_grad_finalize_kernel<<<cdiv((long)b * C_cond, 256), 256, 0, st>>>(S0, S1, a, s, mean, inv_std, G, b,
                                                                         C_cond, stride); count_launch();
}

// ---- transposed GEMV: dcond[bi][k] = sum_j G[bi][j] * W[j][k] (J ~ 10^4..10^5 rows, cdim = 256 columns).
// Pass 1: block = kDcRows rows of W, thread = 4 consecutive k (one 16-byte load per row), 24 x 4
// accumulators, G slice staged transposed in shared memory (same inner loop as gemv_fwd_kernel);
// partial sums go to partial[block][bi][k] — no atomics. Pass 2 (dcond_reduce_split_kernel) sums the
// partials of both operands and writes dz / dc.
constexpr int kDcRows = 128;
__global__ void __launch_bounds__(64) dcond_partial_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ W,
                                                          float* __restrict__ partial, int b, int J, int cdim) {
    extern __shared__ float4 gv_smem[];
    float* gs = reinterpret_cast<float*>(gv_smem);
    const int j0 = blockIdx.x * kDcRows;
    const int nj = min(kDcRows, J - j0);
    const int b0 = blockIdx.y * kGvB;
    for (int i = threadIdx.x; i < kDcRows * kGvB; i += blockDim.x) {
        const int bi = i / kDcRows, jj = i - bi * kDcRows;   // coalesced along j
        gs[jj * kGvB + bi] = (jj < nj && b0 + bi < b) ? G[(long)(b0 + bi) * ldg + j0 + jj] : 0.f;
    }
    __syncthreads();
    const int k = threadIdx.x * 4;
    float acc[kGvB][4];
#pragma unroll
    for (int i = 0; i < kGvB; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    const float* wp = W + (long)j0 * cdim + k;
    for (int jj = 0; jj < nj; jj += kGvU) {
        float4 w[kGvU];
#pragma unroll
        for (int u = 0; u < kGvU; ++u)
            w[u] = (jj + u < nj) ? __ldg(reinterpret_cast<const float4*>(wp + (long)(jj + u) * cdim)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < kGvU; ++u) gv_fma(acc, gv_smem, jj + u, w[u]);   // rows beyond nj are zero in gs
    }
    float* dst = partial + ((long)blockIdx.x * b + b0) * cdim + k;
#pragma unroll
    for (int i = 0; i < kGvB; ++i)
        if (b0 + i < b) *reinterpret_cast<float4*>(dst + (long)i * cdim) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
}
int k_dcond_blocks(int J) { return cdiv(J, kDcRows); }
void k_dcond_partial(const float* G, int ldg, const float* W, float* partial, int b, int J, int cdim, cudaStream_t st) {
    const dim3 grid(cdiv(J, kDcRows), cdiv(b, kGvB));
    dcond_partial_kernel<<<grid, cdim / 4, (size_t)kDcRows * kGvB * sizeof(float), st>>>(G, ldg, W, partial, b, J, cdim);
    count_launch();
}
// dcond = sum over nblk partial blocks; split into dz | dc with the step's gradient scale. block = (64 k, 8 slices)
__global__ void dcond_reduce_split_kernel(const float* __restrict__ partial, int nblk, float* dz, float* dc, int b, int zd,
                                          int cd, float scale, const float* row_scale) {
    __shared__ float red[8][64];
    const int D = zd + cd;
    const int i = blockIdx.x * 64 + threadIdx.x;   // flat (bi, k)
    float acc = 0.f;
    if (i < b * D)
        for (int blk = threadIdx.y; blk < nblk; blk += 8) acc += __ldg(partial + (long)blk * b * D + i);
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y != 0 || i >= b * D) return;
#pragma unroll
    for (int y = 1; y < 8; ++y) acc += red[y][threadIdx.x];
    const int bi = i / D, k = i - bi * D;
    const float sc = row_scale ? scale * row_scale[bi] : scale;
    if (k < zd) dz[bi * zd + k] = acc * sc;
    else dc[bi * cd + k - zd] = acc * sc;
}
void k_dcond_reduce_split(const float* partial, int nblk, float* dz, float* dc, int b, int zd, int cd, float scale,
                          const float* row_scale, cudaStream_t st) {
    dcond_reduce_split_kernel<<<cdiv(b * (zd + cd), 64), dim3(64, 8), 0, st>>>(partial, nblk, dz, dc, b, zd, cd, scale, row_scale);
    count_launch();
}

__global__ void bf16_to_f32_kernel(const bf16* src, float* dst, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = b2f(src[i]);
}
void k_bf16_to_f32(const bf16* src, float* dst, long n, cudaStream_t st) {
    bf16_to_f32_kernel<<<cdiv(n, 256), 256, 0, st>>>(src, dst, n); count_launch();
}

__global__ void split_dcond_kernel(const float* dcond, float* dz, float* dc, int b, int zd, int cd, float scale,
                                   const float* row_scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int D = zd + cd;
    if (i >= b * D) return;
    const int bi = i / D, k = i % D;
    const float sc = row_scale ? scale * row_scale[bi] : scale;
    if (k < zd) dz[bi * zd + k] = dcond[i] * sc;
    else dc[bi * cd + k - zd] = dcond[i] * sc;
}
void k_split_dcond(const float* dcond, float* dz, float* dc, int b, int zd, int cd, float scale,
                   const float* row_scale, cudaStream_t st) {
    split_dcond_kernel<<<cdiv((long)b * (zd + cd), 256), 256, 0, st>>>(dcond, dz, dc, b, zd, cd, scale, row_scale); count_launch();
}

__global__ void fill_kernel(float* p, float v, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
void k_fill_f32(float* p, float v, long n, cudaStream_t st) { fill_kernel<<<cdiv(n, 256), 256, 0, st>>>(p, v, n); count_launch(); }

// ============================================================================= BigGAN glue
__global__ void rgb_gather_kernel(const float* __restrict__ T, const float* __restrict__ bias, float* __restrict__ img,
                                  int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const long plane = (long)H * W;
    const float* Tn = T + (long)n * 27 * plane;
    float acc[3] = {__ldg(bias), __ldg(bias + 1), __ldg(bias + 2)};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int yy = y + r - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int xx = x + s - 1;
            if (xx < 0 || xx >= W) continue;
            const float* t = Tn + (long)((r * 3 + s) * 3) * plane + (long)yy * W + xx;
            acc[0] += __ldg(t);
            acc[1] += __ldg(t + plane);
            acc[2] += __ldg(t + 2 * plane);
        }
    }
#pragma unroll
    for (int o = 0; o < 3; ++o) img[((long)n * 3 + o) * plane + (long)y * W + x] = tanhf(acc[o]);
}
void k_rgb_gather(const float* T, const float* bias, float* img, int b, int H, int W, cudaStream_t st) {
    rgb_gather_kernel<<<dim3(cdiv(W, 128), H, b), 128, 0, st>>>(T, bias, img, H, W); count_launch();
}

// thread = 8 channels (one uint4) of one low-res pixel; block = 8 channel-groups x 32 pixels
__device__ __forceinline__ void unpack8(const uint4 t, float (&f)[8]) {
    f[0] = act_lo(t.x); f[1] = act_hi(t.x); f[2] = act_lo(t.y); f[3] = act_hi(t.y);
    f[4] = act_lo(t.z); f[5] = act_hi(t.z); f[6] = act_lo(t.w); f[7] = act_hi(t.w);
}
__device__ __forceinline__ uint32_t pk2(float a, float b) { return pack_act(a, b); }
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pk2(f[0], f[1]), pk2(f[2], f[3]), pk2(f[4], f[5]), pk2(f[6], f[7]));
}
__global__ void pool_bnrelu_bwd_kernel(const bf16* __restrict__ g_up, const bf16* __restrict__ y_lo,
                                       const float* __restrict__ a, int aff_stride, float* __restrict__ statp,
                                       int parts, bf16* dx, int H, int W, int C) {
    __shared__ float r0[32][65], r1[32][65];
    const int cg = threadIdx.x, py = threadIdx.y;  // 8 x 32
    const int c = blockIdx.y * 64 + cg * 8;
    const int bi = blockIdx.z;
    const int HW = H * W;
    const int W2 = 2 * W;
    float av[8], s0[8], s1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { av[i] = a[(long)bi * aff_stride + c + i]; s0[i] = 0.f; s1[i] = 0.f; }
    for (int p = blockIdx.x * 256 + py; p < min(HW, (int)(blockIdx.x + 1) * 256); p += 32) {
        const int y = p / W, x = p % W;
        const long base = (((long)bi * 2 * H + 2 * y) * W2 + 2 * x) * C + c;
        float g[8], t[8], yv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(g_up + base)), g);
        unpack8(__ldg(reinterpret_cast<const uint4*>(g_up + base + C)), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] += t[i];
        unpack8(__ldg(reinterpret_cast<const uint4*>(g_up + base + (long)W2 * C)), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] += t[i];
        unpack8(__ldg(reinterpret_cast<const uint4*>(g_up + base + (long)W2 * C + C)), t);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] += t[i];
        const long o = ((long)bi * HW + p) * C + c;
        unpack8(__ldg(reinterpret_cast<const uint4*>(y_lo + o)), yv);
        float d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float dpre = yv[i] > 0.f ? g[i] : 0.f;
            s0[i] += dpre;
            s1[i] += dpre * yv[i];
            d[i] = av[i] * dpre;
        }
        *reinterpret_cast<uint4*>(dx + o) = pack8(d);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { r0[py][cg * 8 + i] = s0[i]; r1[py][cg * 8 + i] = s1[i]; }
    __syncthreads();
    const int tid = py * 8 + cg;
    if (tid < 64) {
        float t0 = 0.f, t1 = 0.f;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) { t0 += r0[k][tid]; t1 += r1[k][tid]; }
        // one writer per (image, pixel block, channel): the fixed-order sum over the blocks follows in stat_reduce_kernel
        float* dst = statp + ((long)bi * parts + blockIdx.x) * 2 * C + blockIdx.y * 64 + tid;
        dst[0] = t0;
        dst[C] = t1;
    }
}
int k_pool_bnrelu_parts(int H, int W) { return cdiv((long)H * W, 256); }
void k_pool_bnrelu_bwd(const bf16* g_up, const bf16* y_lo, const float* a, int aff_stride, float* statp,
                       bf16* dx, int b, int H, int W, int C, cudaStream_t st) {
    dim3 grid(cdiv((long)H * W, 256), C / 64, b), block(8, 32);
    pool_bnrelu_bwd_kernel<<<grid, block, 0, st>>>(g_up, y_lo, a, aff_stride, statp, (int)grid.x, dx, H, W, C); count_launch();
}

// BN-gradient sums: S0/S1[n][off + c] = sum over the layer's partial slots in a fixed order: 32 interleaved slices of the slot
// sequence per (n, c) — warp w keeps the running sums of slices w, w+8, w+16, w+24, i.e. eight independent loads in flight per
// thread — then a fixed tree over the 32 slice sums. The order depends on the number of slots only, never on the launch
// geometry of the producers or of this kernel.
__global__ void __launch_bounds__(256) stat_reduce_kernel(const StatSeg* __restrict__ segs, float* __restrict__ S0,
                                                          float* __restrict__ S1, int stride, int stride1) {
    __shared__ float r0[32][33], r1[32][33];
    const StatSeg sg = segs[blockIdx.x];
    const int n = blockIdx.y, lane = threadIdx.x, w = threadIdx.y;
    const float* base = sg.p + (long)n * sg.pstride * 2 * sg.C + sg.c0 + lane;
    const long step = 2L * sg.C;
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
    for (int q = w; q < sg.parts; q += 32) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (q + 8 * u < sg.parts) {
                a0[u] += __ldg(base + (q + 8 * u) * step);
                a1[u] += __ldg(base + (q + 8 * u) * step + sg.C);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { r0[w + 8 * u][lane] = a0[u]; r1[w + 8 * u][lane] = a1[u]; }
    __syncthreads();
    if (w == 0) {
        float t0[8], t1[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            t0[k] = (r0[4 * k][lane] + r0[4 * k + 1][lane]) + (r0[4 * k + 2][lane] + r0[4 * k + 3][lane]);
            t1[k] = (r1[4 * k][lane] + r1[4 * k + 1][lane]) + (r1[4 * k + 2][lane] + r1[4 * k + 3][lane]);
        }
        S0[(long)n * stride + sg.off + sg.c0 + lane] = ((t0[0] + t0[1]) + (t0[2] + t0[3])) + ((t0[4] + t0[5]) + (t0[6] + t0[7]));
        S1[(long)n * stride1 + sg.off1 + sg.c0 + lane] = ((t1[0] + t1[1]) + (t1[2] + t1[3])) + ((t1[4] + t1[5]) + (t1[6] + t1[7]));
    }
}
void k_stat_reduce(const StatSeg* segs, int nsegs, float* S0, float* S1, int stride, int b, cudaStream_t st) {
    if (nsegs <= 0) return;
    stat_reduce_kernel<<<dim3(nsegs, b), dim3(32, 8), 0, st>>>(segs, S0, S1, stride, stride); count_launch();
}
void k_stat_reduce2(const StatSeg* segs, int nsegs, float* S0, int stride0, float* S1, int stride1, int b, cudaStream_t st) {
    if (nsegs <= 0) return;
    stat_reduce_kernel<<<dim3(nsegs, b), dim3(32, 8), 0, st>>>(segs, S0, S1, stride0, stride1); count_launch();
}

__global__ void pool2x2_sum_kernel(const bf16* __restrict__ in, int inC, bf16* __restrict__ out, int b, int H, int W, int C) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int CG = C / 8;
    const long total = (long)b * H * W * CG;
    if (i >= total) return;
    const int cg = i % CG;
    const long q = i / CG;
    const int x = q % W, y = (q / W) % H, bi = q / ((long)W * H);
    const long W2 = 2L * W;
    const long base = (((long)bi * 2 * H + 2 * y) * W2 + 2 * x) * inC + cg * 8;
    float a[8], t[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(in + base)), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(in + base + inC)), t);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += t[e];
    unpack8(__ldg(reinterpret_cast<const uint4*>(in + base + W2 * inC)), t);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += t[e];
    unpack8(__ldg(reinterpret_cast<const uint4*>(in + base + W2 * inC + inC)), t);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += t[e];
    *reinterpret_cast<uint4*>(out + q * C + cg * 8) = pack8(a);
}
void k_pool2x2_sum(const bf16* in, int inC, bf16* out, int b, int H, int W, int C, cudaStream_t st) {
    const long total = (long)b * H * W * (C / 8);
    pool2x2_sum_kernel<<<cdiv(total, 256), 256, 0, st>>>(in, inC, out, b, H, W, C); count_launch();
}

// ============================================================================= attention glue
// 2x2 max-pool of a channel slice, vectorised: block = 32 pooled pixels x 64 channels (thread = 8 channels of one pooled
// pixel: four 16-byte loads, one 16-byte store of `out`, 8 argmax bytes); the transposed copy leaves through a shared-memory
// tile so that its rows (one channel, 32 consecutive pooled pixels) are written as contiguous 64-byte runs
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const bf16* __restrict__ x, int xC, int c0, int C, bf16* __restrict__ out,
                                                           bf16* __restrict__ outT, unsigned char* __restrict__ idx, int H, int W) {
    __shared__ bf16 tile[64][40];   // [channel][pooled pixel] (+8 padding: 80-byte rows keep 16-byte alignment)
    const int Ho = H / 2, Wo = W / 2, nk = Ho * Wo;
    const int bi = blockIdx.z, cb = blockIdx.y * 64, k0 = blockIdx.x * 32;
    const int cg = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const int kk = k0 + pl;
    if (kk < nk) {
        const int oy = kk / Wo, ox = kk - oy * Wo;
        float best[8];
        int bidx[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int y = 2 * oy + (k >> 1), xx = 2 * ox + (k & 1);
            float v[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long)bi * H + y) * W + xx) * xC + c0 + cb + cg * 8)), v);
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (k == 0 || v[e] > best[e]) { best[e] = v[e]; bidx[e] = k; }
        }
        const long o = ((long)bi * nk + kk) * C + cb + cg * 8;
        const uint4 pk = pack8(best);
        if (out) *reinterpret_cast<uint4*>(out + o) = pk;
        uint2 ib;
        ib.x = bidx[0] | (bidx[1] << 8) | (bidx[2] << 16) | (bidx[3] << 24);
        ib.y = bidx[4] | (bidx[5] << 8) | (bidx[6] << 16) | (bidx[7] << 24);
        *reinterpret_cast<uint2*>(idx + o) = ib;
        const bf16* h = reinterpret_cast<const bf16*>(&pk);
#pragma unroll
        for (int e = 0; e < 8; ++e) tile[cg * 8 + e][pl] = h[e];
    }
    __syncthreads();
    if (outT) {
        // thread = (channel, quarter of the 32 pixels): one 16-byte store
        const int c = threadIdx.x >> 2, q = threadIdx.x & 3;
        if (k0 + q * 8 < nk)
            *reinterpret_cast<uint4*>(outT + ((long)bi * C + cb + c) * nk + k0 + q * 8) = *reinterpret_cast<const uint4*>(&tile[c][q * 8]);
    }
}
void k_maxpool2_fwd(const bf16* x, int xC, int c0, int C, bf16* out, bf16* outT, unsigned char* idx, int b,
                    int H, int W, cudaStream_t st) {
    // C % 64 == 0, (H/2 * W/2) % 8 == 0 (the attention shapes: 64 / 256 channels, >= 128 pooled pixels)
    const int nk = (H / 2) * (W / 2);
    maxpool2_fwd_kernel<<<dim3(cdiv(nk, 32), C / 64, b), 256, 0, st>>>(x, xC, c0, C, out, outT, idx, H, W); count_launch();
}

// thread = 8 channels of one pooled pixel: routes the gradient to the argmax position, zeros to the other three
__global__ void maxpool2_bwd_kernel(const bf16* __restrict__ d_out, const unsigned char* __restrict__ idx,
                                    bf16* __restrict__ dx, int xC, int c0, int C, int b, int H, int W) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int Ho = H / 2, Wo = W / 2, CG = C / 8;
    const long total = (long)b * Ho * Wo * CG;
    if (i >= total) return;
    const int cg = i % CG;
    const long q = i / CG;
    const int ox = q % Wo, oy = (q / Wo) % Ho, bi = q / ((long)Wo * Ho);
    const long o = q * C + cg * 8;
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(d_out + o));
    const uint2 ib = __ldg(reinterpret_cast<const uint2*>(idx + o));
    const unsigned short* gh = reinterpret_cast<const unsigned short*>(&g);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int y = 2 * oy + (k >> 1), xx = 2 * ox + (k & 1);
        unsigned short r[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int id = ((e < 4 ? ib.x : ib.y) >> ((e & 3) * 8)) & 0xFF;
            r[e] = (id == k) ? gh[e] : (unsigned short)0;
        }
        *reinterpret_cast<uint4*>(dx + (((long)bi * H + y) * W + xx) * xC + c0 + cg * 8) = *reinterpret_cast<const uint4*>(r);
    }
}
void k_maxpool2_bwd(const bf16* d_out, const unsigned char* idx, bf16* dx, int xC, int c0, int C, int b, int H,
                    int W, cudaStream_t st) {
    const long total = (long)b * (H / 2) * (W / 2) * (C / 8);
    maxpool2_bwd_kernel<<<cdiv(total, 256), 256, 0, st>>>(d_out, idx, dx, xC, c0, C, b, H, W); count_launch();
}

// one warp per row
__global__ void softmax_fwd_kernel(const float* __restrict__ S, bf16* __restrict__ P, long rows, int n) {
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* s = S + row * n;
    float m = -INFINITY;
    for (int k = lane; k < n; k += 32) m = fmaxf(m, s[k]);
    m = warp_max(m);
    float sum = 0.f;
    for (int k = lane; k < n; k += 32) sum += __expf(s[k] - m);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    bf16* p = P + row * n;
    for (int k = lane; k < n; k += 32) p[k] = f2b(__expf(s[k] - m) * inv);
}
void k_softmax_fwd(const float* S, bf16* P, long rows, int n, cudaStream_t st) {
    softmax_fwd_kernel<<<cdiv(rows, 8), 256, 0, st>>>(S, P, rows, n); count_launch();
}

__global__ void softmax_bwd_kernel(const bf16* __restrict__ P, const float* __restrict__ dP, bf16* __restrict__ dS,
                                   long rows, int n) {
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const bf16* p = P + row * n;
    const float* dp = dP + row * n;
    float dot = 0.f;
    for (int k = lane; k < n; k += 32) dot = fmaf(dp[k], b2f(p[k]), dot);
    dot = warp_sum(dot);
    bf16* ds = dS + row * n;
    for (int k = lane; k < n; k += 32) ds[k] = f2b(b2f(p[k]) * (dp[k] - dot));
}
void k_softmax_bwd(const bf16* P, const float* dP, bf16* dS, long rows, int n, cudaStream_t st) {
    softmax_bwd_kernel<<<cdiv(rows, 8), 256, 0, st>>>(P, dP, dS, rows, n); count_launch();
}

// one warp per row, 16-byte loads
__global__ void rowdot_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, float* __restrict__ D, long rows, int C) {
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float acc = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
        const uint4 ua = __ldg(reinterpret_cast<const uint4*>(a + row * C + c));
        const uint4 ub = __ldg(reinterpret_cast<const uint4*>(b + row * C + c));
        acc += act_lo(ua.x) * act_lo(ub.x) + act_hi(ua.x) * act_hi(ub.x) + act_lo(ua.y) * act_lo(ub.y) + act_hi(ua.y) * act_hi(ub.y) +
               act_lo(ua.z) * act_lo(ub.z) + act_hi(ua.z) * act_hi(ub.z) + act_lo(ua.w) * act_lo(ub.w) + act_hi(ua.w) * act_hi(ub.w);
    }
    acc = warp_sum(acc);
    if (lane == 0) D[row] = acc;
}
void k_rowdot(const bf16* a, const bf16* b, float* D, long rows, int C, cudaStream_t st) {
    rowdot_kernel<<<cdiv(rows, 8), 256, 0, st>>>(a, b, D, rows, C); count_launch();
}

__global__ void transpose_kernel(const bf16* __restrict__ in, int ldin, int in_c0, bf16* __restrict__ out, int R, int C) {
    // 64 x 64 tile; rows padded to 72 elements so column reads spread over the banks
    __shared__ __align__(16) bf16 tile[64][72];
    const int bi = blockIdx.z;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    const bf16* src = in + (long)bi * R * ldin + in_c0;
    bf16* dst = out + (long)bi * C * R;
    const int t = threadIdx.x;  // 256 threads
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = t + i * 256;       // 512 pieces of 8 elements
        const int r = idx >> 3, cp = (idx & 7) * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r0 + r < R && c0 + cp < C) v = __ldg(reinterpret_cast<const uint4*>(src + (long)(r0 + r) * ldin + c0 + cp));
        *reinterpret_cast<uint4*>(&tile[r][cp]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = t + i * 256;
        const int c = idx >> 3, rp = (idx & 7) * 8;
        if (c0 + c < C && r0 + rp < R) {
            __align__(16) bf16 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = tile[rp + k][c];
            *reinterpret_cast<uint4*>(dst + (long)(c0 + c) * R + r0 + rp) = *reinterpret_cast<const uint4*>(v);
        }
    }
}
void k_transpose(const bf16* in, int ldin, int in_c0, bf16* out, int b, int R, int C, cudaStream_t st) {
    // requires R % 8 == 0, C % 8 == 0, 16-byte aligned rows (true for every caller)
    dim3 grid(cdiv(C, 64), cdiv(R, 64), b);
    transpose_kernel<<<grid, 256, 0, st>>>(in, ldin, in_c0, out, R, C); count_launch();
}

// ============================================================================= image / loss glue
__constant__ float kLpipsShift[3] = {-.030f, -.088f, -.188f};
__constant__ float kLpipsScale[3] = {.458f, .448f, .450f};

// AlexNet conv1 (11x11, stride 4, pad 2) as a GEMM: col[(bi, oy, ox)][k], k = (c*11 + r)*11 + s (363 of Kp = 384 columns).
// block = one output row oy x 32 consecutive ox of one image: the (3 x 11 x 135)-pixel input patch is read once, coalesced,
// normalised (LPIPS ScalingLayer) into shared memory; thread (kg, pl) then owns the 8 columns k = 8 kg .. 8 kg + 7 — their
// patch offsets are computed once and live in registers — and writes one 16-byte store per pixel (pixels pl, pl + 8, ...):
// the 48 threads of a pixel cover its 768 contiguous bytes.
constexpr int kA1Pix = 32, kA1PW = 4 * (kA1Pix - 1) + 11;   // 135 patch columns
__global__ void __launch_bounds__(384) im2col_alex1_kernel(const float* __restrict__ img, bf16* __restrict__ col, int H, int W,
                                                           int Ho, int Wo, int Kp) {
    __shared__ float patch[33][kA1PW + 1];   // [(c, r)][column]
    const int bi = blockIdx.z, oy = blockIdx.y, ox0 = blockIdx.x * kA1Pix;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = 4 * ox0 - 2, y0 = 4 * oy - 2;
    for (int cr = warp; cr < 33; cr += 12) {
        const int c = cr / 11, r = cr - c * 11, y = y0 + r;
        const float sh = kLpipsShift[c];
        const float* row = img + (((long)bi * 3 + c) * H + y) * W;
        for (int j = lane; j < kA1PW; j += 32) {
            const int x = x0 + j;
            // (v - shift) / scale as the per-element kernel computed it
            patch[cr][j] = (y >= 0 && y < H && x >= 0 && x < W) ? (__ldg(row + x) - sh) / kLpipsScale[c] : 0.f;
        }
    }
    __syncthreads();
    const int kg = threadIdx.x % 48, pl = threadIdx.x / 48;   // 48 column groups x 8 pixel lanes
    int off[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = kg * 8 + e;
        const int cr = k / 11, s = k - cr * 11;
        off[e] = k < 363 ? cr * (kA1PW + 1) + s : -1;
    }
    const float* pf = &patch[0][0];
    for (int p = pl; p < kA1Pix; p += 8) {
        const int ox = ox0 + p;
        if (ox >= Wo) break;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = off[e] >= 0 ? pf[off[e] + 4 * p] : 0.f;
        *reinterpret_cast<uint4*>(col + (((long)bi * Ho + oy) * Wo + ox) * Kp + kg * 8) = pack8(v);
    }
}
void k_im2col_alex1(const float* img, bf16* col, int b, int H, int W, int Ho, int Wo, int Kp, cudaStream_t st) {
    if (Kp != 384) { fprintf(stderr, "[p2l] im2col_alex1: Kp = %d (expected 384)\n", Kp); abort(); }
    im2col_alex1_kernel<<<dim3(cdiv(Wo, kA1Pix), Ho, b), 384, 0, st>>>(img, col, H, W, Ho, Wo, Kp); count_launch();
}

__global__ void col2im_alex1_kernel(const bf16* __restrict__ dcol, float* __restrict__ dimg, int b, int H, int W,
                                    int Ho, int Wo, int Kp, int accumulate, float unscale) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)b * 3 * H * W;
    if (i >= total) return;
    const int x = i % W, y = (i / W) % H, c = (i / ((long)W * H)) % 3, bi = i / ((long)3 * W * H);
    // the (up to) 3 x 3 output pixels whose 11 x 11 windows cover (y, x): oy = (y+2)/4 - {2,1,0} with r = (y+2)%4 + {8,4,0}.
    // Branch-free (clamped address, contribution zeroed) so that the nine loads are in flight together; same sum order
    // (oy ascending, then ox ascending) as the bounded loops.
    const int yy = y + 2, xx = x + 2;
    const int oyh = yy >> 2, r0 = yy & 3, oxh = xx >> 2, s0 = xx & 3;
    float t[9];
    bool ok[9];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int oy = oyh - 2 + a, r = r0 + 8 - 4 * a;
        const bool vy = oy >= 0 && oy < Ho && r <= 10;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int ox = oxh - 2 + d, s2 = s0 + 8 - 4 * d;
            const bool v = vy && ox >= 0 && ox < Wo && s2 <= 10;
            const long idx = v ? ((((long)bi * Ho + oy) * Wo + ox) * Kp + (c * 11 + r) * 11 + s2) : 0;
            ok[a * 3 + d] = v;
            t[a * 3 + d] = b2f(dcol[idx]);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 9; ++q) acc += ok[q] ? t[q] : 0.f;
    acc *= unscale / kLpipsScale[c];
    dimg[i] = accumulate ? dimg[i] + acc : acc;
}
void k_col2im_alex1(const bf16* dcol, float* dimg, int b, int H, int W, int Ho, int Wo, int Kp, int accumulate,
                    float unscale, cudaStream_t st) {
    const long total = (long)b * 3 * H * W;
    col2im_alex1_kernel<<<cdiv(total, 256), 256, 0, st>>>(dcol, dimg, b, H, W, Ho, Wo, Kp, accumulate, unscale); count_launch();
}

__global__ void maxpool_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, unsigned char* __restrict__ idx,
                                   int b, int H, int W, int C, int Ho, int Wo, int k, int s) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int CG = C / 8;
    const long total = (long)b * Ho * Wo * CG;
    if (i >= total) return;
    const int cg = i % CG;
    const long q = i / CG;
    const int ox = q % Wo, oy = (q / Wo) % Ho, bi = q / ((long)Wo * Ho);
    float best[8];
    int bidx[8];
    for (int r = 0; r < k; ++r) {
        for (int t = 0; t < k; ++t) {
            float v[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(x + (((long)bi * H + oy * s + r) * W + ox * s + t) * C + cg * 8)), v);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if ((r == 0 && t == 0) || v[e] > best[e]) { best[e] = v[e]; bidx[e] = r * k + t; }
            }
        }
    }
    *reinterpret_cast<uint4*>(out + q * C + cg * 8) = pack8(best);
    uint2 pk;
    pk.x = bidx[0] | (bidx[1] << 8) | (bidx[2] << 16) | (bidx[3] << 24);
    pk.y = bidx[4] | (bidx[5] << 8) | (bidx[6] << 16) | (bidx[7] << 24);
    *reinterpret_cast<uint2*>(idx + q * C + cg * 8) = pk;
}
void k_maxpool_fwd(const bf16* x, bf16* out, unsigned char* idx, int b, int H, int W, int C, int Ho, int Wo, int k,
                   int s, cudaStream_t st) {
    const long total = (long)b * Ho * Wo * (C / 8);
    maxpool_fwd_kernel<<<cdiv(total, 256), 256, 0, st>>>(x, out, idx, b, H, W, C, Ho, Wo, k, s); count_launch();
}

__global__ void maxpool_bwd_kernel(const bf16* __restrict__ dout, const unsigned char* __restrict__ idx,
                                   const bf16* __restrict__ x, const bf16* __restrict__ addin, bf16* __restrict__ dx,
                                   int b, int H, int W, int C, int Ho, int Wo, int k, int s) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int CG = C / 8;
    const long total = (long)b * H * W * CG;
    if (i >= total) return;
    const int cg = i % CG;
    const long q = i / CG;
    const int xx = q % W, y = (q / W) % H, bi = q / ((long)W * H);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    // windows (oy, ox) with oy*s <= y < oy*s + k
    const int oy_lo = max(0, (y - k + s) / s), oy_hi = min(Ho - 1, y / s);
    const int ox_lo = max(0, (xx - k + s) / s), ox_hi = min(Wo - 1, xx / s);
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
        const int r = y - oy * s;
        if (r < 0 || r >= k) continue;
        for (int ox = ox_lo; ox <= ox_hi; ++ox) {
            const int t = xx - ox * s;
            if (t < 0 || t >= k) continue;
            const long o = (((long)bi * Ho + oy) * Wo + ox) * C + cg * 8;
            const uint2 pk = __ldg(reinterpret_cast<const uint2*>(idx + o));
            float g[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(dout + o)), g);
            const int me = r * k + t;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int id = ((e < 4 ? pk.x : pk.y) >> ((e & 3) * 8)) & 0xFF;
                if (id == me) acc[e] += g[e];
            }
        }
    }
    const long o = q * C + cg * 8;
    if (x) {
        float xv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + o)), xv);
#pragma unroll
        for (int e = 0; e < 8; ++e) if (!(xv[e] > 0.f)) acc[e] = 0.f;
    }
    if (addin) {
        float av[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(addin + o)), av);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += av[e];
    }
    *reinterpret_cast<uint4*>(dx + o) = pack8(acc);
}
void k_maxpool_bwd(const bf16* dout, const unsigned char* idx, const bf16* x, const bf16* addin, bf16* dx, int b,
                   int H, int W, int C, int Ho, int Wo, int k, int s, cudaStream_t st) {
    const long total = (long)b * H * W * (C / 8);
    maxpool_bwd_kernel<<<cdiv(total, 256), 256, 0, st>>>(dout, idx, x, addin, dx, b, H, W, C, Ho, Wo, k, s); count_launch();
}

// one warp per feature pixel, block = 8 warps; a lane keeps its <= 16 channels of the pixel (features, target, lin
// weights) in registers, so HBM is read ONCE; block k of sample bi writes its part of the loss to slot k
// NC = channels per lane = ceil(C / 32): 2 .. 16 (alex: 64..384 channels, vgg16: 64..512) — an instantiation per width, so
// that the 64-channel first layer (most pixels by far) does not issue sixteen predicated passes
template <int NC>
__global__ void __launch_bounds__(256) lpips_dist_kernel(const bf16* __restrict__ f, const float* __restrict__ t,
                                  const float* __restrict__ lin, const float* __restrict__ wadj, float* __restrict__ lossp,
                                  int lp_stride, bf16* __restrict__ g, int HW, int C, float gscale) {
    __shared__ float part[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bi = blockIdx.y;
    const int p = blockIdx.x * 8 + warp;
    float contrib = 0.f;
    if (p < HW) {
        const bf16* fp = f + ((long)bi * HW + p) * C;
        const float* tp = t + (long)p * C;
        float v[NC], tv[NC], lv[NC];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const int c = lane + 32 * i;
            const bool in = c < C;
            v[i] = in ? b2f(fp[c]) : 0.f;
            tv[i] = in ? __ldg(tp + c) : 0.f;
            lv[i] = in ? __ldg(lin + c) : 0.f;
            ss = fmaf(v[i], v[i], ss);
        }
        ss = warp_sum(ss);
        const float r = sqrtf(ss);
        const float inv = 1.f / (r + 1e-10f);
        float d = 0.f, q = 0.f;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const float diff = v[i] * inv - tv[i];
            d = fmaf(lv[i] * diff, diff, d);
            q = fmaf(2.f * lv[i] * diff, v[i], q);
        }
        d = warp_sum(d);
        q = warp_sum(q);
        const float wv = wadj[p];
        contrib = wv * d;
        if (g) {
            bf16* gp = g + ((long)bi * HW + p) * C;
            const float k2 = r > 0.f ? q * inv * inv / r : 0.f;
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const int c = lane + 32 * i;
                if (c < C) {
                    const float diff = v[i] * inv - tv[i];
                    const float e = 2.f * lv[i] * diff;
                    const float df = e * inv - k2 * v[i];
                    gp[c] = f2b(v[i] > 0.f ? wv * gscale * df : 0.f);
                }
            }
        }
    }
    if (lane == 0) part[warp] = contrib;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += part[i];
        lossp[(long)bi * lp_stride + blockIdx.x] = s;   // one slot per block: summed in slot order by loss_reduce_kernel
    }
}
int k_lpips_dist_slots(int HW) { return cdiv(HW, 8); }
void k_lpips_dist(const bf16* f, const float* t, const float* lin, const float* wadj, float* lossp, int lp_stride, bf16* g,
                  int b, int HW, int C, float gscale, cudaStream_t st) {
    dim3 grid(cdiv(HW, 8), b);
    const int nc = cdiv(C, 32);
    if (nc <= 2) lpips_dist_kernel<2><<<grid, 256, 0, st>>>(f, t, lin, wadj, lossp, lp_stride, g, HW, C, gscale);
    else if (nc <= 4) lpips_dist_kernel<4><<<grid, 256, 0, st>>>(f, t, lin, wadj, lossp, lp_stride, g, HW, C, gscale);
    else if (nc <= 8) lpips_dist_kernel<8><<<grid, 256, 0, st>>>(f, t, lin, wadj, lossp, lp_stride, g, HW, C, gscale);
    else if (nc <= 12) lpips_dist_kernel<12><<<grid, 256, 0, st>>>(f, t, lin, wadj, lossp, lp_stride, g, HW, C, gscale);
    else if (nc <= 16) lpips_dist_kernel<16><<<grid, 256, 0, st>>>(f, t, lin, wadj, lossp, lp_stride, g, HW, C, gscale);
    else { fprintf(stderr, "[p2l] lpips_dist: %d channels (max 512)\n", C); abort(); }
    count_launch();
}

// loss[bi] = sum of the sample's partial slots (pixel term blocks, then every layer's distance blocks), fixed order
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ lossp, int nslots, int lp_stride,
                                                          float* __restrict__ loss) {
    __shared__ float red[256];
    const int bi = blockIdx.x, t = threadIdx.x;
    float a = 0.f;
    for (int k = t; k < nslots; k += 256) a += lossp[(long)bi * lp_stride + k];
    red[t] = a;
    __syncthreads();
    for (int h = 128; h >= 1; h >>= 1) {
        if (t < h) red[t] += red[t + h];
        __syncthreads();
    }
    if (t == 0) loss[bi] = red[0];
}
void k_loss_reduce(const float* lossp, int nslots, int lp_stride, float* loss, int b, cudaStream_t st) {
    loss_reduce_kernel<<<b, 256, 0, st>>>(lossp, nslots, lp_stride, loss); count_launch();
}

__global__ void lpips_normalize_kernel(const bf16* __restrict__ f, float* __restrict__ t, int HW, int C) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * 8 + warp;
    if (p >= HW) return;
    const bf16* fp = f + (long)p * C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float v = b2f(fp[c]);
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float inv = 1.f / (sqrtf(ss) + 1e-10f);
    for (int c = lane; c < C; c += 32) t[(long)p * C + c] = b2f(fp[c]) * inv;
}
void k_lpips_normalize(const bf16* f, float* t, int HW, int C, cudaStream_t st) {
    lpips_normalize_kernel<<<cdiv(HW, 8), 256, 0, st>>>(f, t, HW, C); count_launch();
}

__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
    float src = scale * (dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 < in_size - 1 ? i0 + 1 : i0;
    l1 = src - i0;
}
// adjoint of the bilinear up-sampling as a GATHER (one thread per low-res pixel sums, in raster order, the hi-res
// pixels whose footprint touches it): no atomics, so two targets built from the same data are bitwise identical
__global__ void upsample_adjoint_kernel(const float* __restrict__ wsum, float* __restrict__ wadj, int H, int W, int h, int w,
                                        float coef) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h * w) return;
    const int y = i / w, x = i % w;
    // hi-res rows whose source coordinate lies in (y - 1, y + 1): src = (h/H)(Y + .5) - .5
    const float sy = (float)H / h, sx = (float)W / w;
    const int Y0 = max(0, (int)floorf((y - 1 + 0.5f) * sy - 0.5f) - 1), Y1 = min(H - 1, (int)ceilf((y + 1 + 0.5f) * sy - 0.5f) + 1);
    const int X0 = max(0, (int)floorf((x - 1 + 0.5f) * sx - 0.5f) - 1), X1 = min(W - 1, (int)ceilf((x + 1 + 0.5f) * sx - 0.5f) + 1);
    float acc = 0.f;
    for (int Y = Y0; Y <= Y1; ++Y) {
        int y0, y1;
        float ly;
        bilinear_src(Y, (float)h / H, h, y0, y1, ly);
        float wy = 0.f;
        if (y0 == y) wy += 1.f - ly;
        if (y1 == y) wy += ly;
        if (wy == 0.f) continue;
        for (int X = X0; X <= X1; ++X) {
            int x0, x1;
            float lx;
            bilinear_src(X, (float)w / W, w, x0, x1, lx);
            float wx = 0.f;
            if (x0 == x) wx += 1.f - lx;
            if (x1 == x) wx += lx;
            if (wx != 0.f) acc += wsum[Y * W + X] * coef * wy * wx;
        }
    }
    wadj[i] = acc;
}
void k_upsample_adjoint(const float* wsum, float* wadj, int H, int W, int h, int w, float coef, cudaStream_t st) {
    upsample_adjoint_kernel<<<cdiv((long)h * w, 64), 64, 0, st>>>(wsum, wadj, H, W, h, w, coef); count_launch();
}

__global__ void __launch_bounds__(1024) weight_sum_kernel(const float* __restrict__ weight, const float* __restrict__ mask,
                                                          float* wsum, float* total, int HW) {
    // ONE block (runs once per target): strided running sums, then a fixed tree — deterministic
    __shared__ float red[1024];
    float acc = 0.f;
    for (int i = threadIdx.x; i < HW; i += 1024) {
        float v = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float wv = weight ? weight[c * HW + i] : 1.f;
            if (mask) wv *= mask[c * HW + i];
            v += wv;
        }
        wsum[i] = v;
        acc += v;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int h = 512; h >= 1; h >>= 1) {
        if ((int)threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
        __syncthreads();
    }
    if (threadIdx.x == 0) total[0] = red[0];
}
void k_weight_sum(const float* weight, const float* mask, float* wsum, float* total, int HW, cudaStream_t st) {
    weight_sum_kernel<<<1, 1024, 0, st>>>(weight, mask, wsum, total, HW); count_launch();
}

// block = 4096 consecutive elements of one sample (16 per thread as four float4); its part of the loss goes to slot
// blockIdx.x of the sample (no atomics)
constexpr int kL1PerBlock = 4096;
__global__ void __launch_bounds__(256) l1_loss_kernel(const float* __restrict__ img, const float* __restrict__ target,
                               const float* __restrict__ weight, const float* __restrict__ mask,
                               const float* __restrict__ total, float* __restrict__ lossp, int lp_stride,
                               float* __restrict__ dimg, int HW3, int l2) {
    __shared__ float part[8];
    const int bi = blockIdx.y;
    const float inv_total = 1.f / total[0];
    const float* im = img + (long)bi * HW3;
    float* dg = dimg ? dimg + (long)bi * HW3 : nullptr;
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = blockIdx.x * kL1PerBlock + (q * 256 + threadIdx.x) * 4;
        if (((HW3 & 3) == 0) && i + 3 < HW3) {   // (odd-sized images: rows of later samples are not 16-byte aligned)
            const float4 o4 = __ldg(reinterpret_cast<const float4*>(im + i));
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(target + i));
            float4 w4 = weight ? __ldg(reinterpret_cast<const float4*>(weight + i)) : make_float4(1.f, 1.f, 1.f, 1.f);
            if (mask) {
                const float4 m4 = __ldg(reinterpret_cast<const float4*>(mask + i));
                w4.x *= m4.x; w4.y *= m4.y; w4.z *= m4.z; w4.w *= m4.w;
            }
            const float o[4] = {o4.x, o4.y, o4.z, o4.w}, tt[4] = {t4.x, t4.y, t4.z, t4.w};
            const float ww[4] = {w4.x * inv_total, w4.y * inv_total, w4.z * inv_total, w4.w * inv_total};
            float gi[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d = tt[e] - o[e];
                if (l2) { v += d * d * ww[e]; gi[e] = -2.f * d * ww[e]; }
                else { v += fabsf(d) * ww[e]; gi[e] = d > 0.f ? -ww[e] : (d < 0.f ? ww[e] : 0.f); }
            }
            if (dg) *reinterpret_cast<float4*>(dg + i) = make_float4(gi[0], gi[1], gi[2], gi[3]);
        } else {
            for (int e = 0; e < 4 && i + e < HW3; ++e) {   // ragged tail (HW3 % 4 != 0)
                float wv = weight ? weight[i + e] : 1.f;
                if (mask) wv *= mask[i + e];
                wv *= inv_total;
                const float d = target[i + e] - im[i + e];
                float gi;
                if (l2) { v += d * d * wv; gi = -2.f * d * wv; }
                else { v += fabsf(d) * wv; gi = d > 0.f ? -wv : (d < 0.f ? wv : 0.f); }
                if (dg) dg[i + e] = gi;
            }
        }
    }
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += part[k];
        lossp[(long)bi * lp_stride + blockIdx.x] = s;
    }
}
int k_l1_loss_slots(int HW3) { return cdiv(HW3, kL1PerBlock); }
void k_l1_loss(const float* img, const float* target, const float* weight, const float* mask, const float* total,
               float* lossp, int lp_stride, float* dimg, int b, int HW3, int HW, int l2, cudaStream_t st) {
    (void)HW;
    dim3 grid(cdiv(HW3, kL1PerBlock), b);
    l1_loss_kernel<<<grid, 256, 0, st>>>(img, target, weight, mask, total, lossp, lp_stride, dimg, HW3, l2); count_launch();
}

__global__ void scale_rows_kernel(float* x, const float* scale, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[(long)blockIdx.y * n + i] *= scale[blockIdx.y];
}
void k_scale_rows(float* x, const float* scale, int b, long n, cudaStream_t st) {
    dim3 grid(cdiv(n, 256), b);
    scale_rows_kernel<<<grid, 256, 0, st>>>(x, scale, n); count_launch();
}

__global__ void im2col_rgb_bwd_kernel(const float* __restrict__ dimg, const float* __restrict__ img,
                                      bf16* __restrict__ col, int b, int H, int W, int Kp, float scale) {
    // thread = one pixel: 27 live values (k = (r*3+s)*3 + o), one full 128-byte row written
    const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)b * H * W;
    if (q >= total) return;
    const int x = q % W, y = (q / W) % H, bi = q / ((long)W * H);
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            // dx[q] = sum_{r,s,o} dv[q - (r-1, s-1), o] * W[o, c, r, s]
            const int yy = y - (r - 1), xx = x - (s - 1);
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
#pragma unroll
                for (int o = 0; o < 3; ++o) {
                    const long oi = (((long)bi * 3 + o) * H + yy) * W + xx;
                    const float im = __ldg(img + oi);
                    v[(r * 3 + s) * 3 + o] = __ldg(dimg + oi) * (1.f - im * im) * scale;
                }
            }
        }
    }
    uint4* dst = reinterpret_cast<uint4*>(col + q * Kp);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        dst[j] = make_uint4(pk2(v[j * 8], v[j * 8 + 1]), pk2(v[j * 8 + 2], v[j * 8 + 3]), pk2(v[j * 8 + 4], v[j * 8 + 5]),
                            pk2(v[j * 8 + 6], v[j * 8 + 7]));
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int j = 4; j < Kp / 8; ++j) dst[j] = z;
}
void k_im2col_rgb_bwd(const float* dimg, const float* img, bf16* col, int b, int H, int W, int Kp, float scale, cudaStream_t st) {
    const long total = (long)b * H * W;
    im2col_rgb_bwd_kernel<<<cdiv(total, 128), 128, 0, st>>>(dimg, img, col, b, H, W, Kp, scale); count_launch();
}

// ---- VGG first layer helpers (3 input channels padded to Cp) --------------------------------
__global__ void img_to_nhwc_scaled_kernel(const float* __restrict__ img, bf16* __restrict__ out, int b, int H, int W,
                                          int Cp) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)b * H * W * Cp;
    if (i >= total) return;
    const int c = i % Cp;
    const long q = i / Cp;
    const int x = q % W, y = (q / W) % H, bi = q / ((long)W * H);
    float v = 0.f;
    if (c < 3) v = (img[(((long)bi * 3 + c) * H + y) * W + x] - kLpipsShift[c]) / kLpipsScale[c];
    out[i] = f2b(v);
}
void k_img_to_nhwc_scaled(const float* img, bf16* out, int b, int H, int W, int Cp, cudaStream_t st) {
    const long total = (long)b * H * W * Cp;
    img_to_nhwc_scaled_kernel<<<cdiv(total, 256), 256, 0, st>>>(img, out, b, H, W, Cp); count_launch();
}
__global__ void nhwc_to_dimg_scaled_kernel(const bf16* __restrict__ dx, int Cp, float* __restrict__ dimg, int b, int H,
                                           int W, int accumulate, float unscale) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)b * 3 * H * W;
    if (i >= total) return;
    const int x = i % W, y = (i / W) % H, c = (i / ((long)W * H)) % 3, bi = i / ((long)3 * W * H);
    const float v = b2f(dx[(((long)bi * H + y) * W + x) * Cp + c]) * unscale / kLpipsScale[c];
    dimg[i] = accumulate ? dimg[i] + v : v;
}
void k_nhwc_to_dimg_scaled(const bf16* dx, int Cp, float* dimg, int b, int H, int W, int accumulate, float unscale, cudaStream_t st) {
    const long total = (long)b * 3 * H * W;
    nhwc_to_dimg_scaled_kernel<<<cdiv(total, 256), 256, 0, st>>>(dx, Cp, dimg, b, H, W, accumulate, unscale); count_launch();
}

// ============================================================================= transform search
// F.affine_grid(theta, size) (align_corners=False): x_j = (2j + 1)/W - 1, y_i = (2i + 1)/H - 1,
//   (gx, gy) = theta[n] . (x_j, y_i, 1)
// F.grid_sample(src, grid) (bilinear, padding zeros, align_corners=False):
//   ix = ((gx + 1) * W - 1) / 2, iy likewise; out = sum over the four neighbours inside the image of
//   src * (1 - |ix - x0|)(1 - |iy - y0|)
__global__ void affine_resample_kernel(const float* __restrict__ src, int src_batch, const float* __restrict__ theta,
                                       float* __restrict__ dst, int b, int C, int H, int W) {
    const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)b * H * W;
    if (q >= total) return;
    const int j = q % W, i = (q / W) % H, n = q / ((long)W * H);
    const float* th = theta + (long)n * 6;
    const float x = (2.f * j + 1.f) / W - 1.f, y = (2.f * i + 1.f) / H - 1.f;
    const float gx = th[0] * x + th[1] * y + th[2];
    const float gy = th[3] * x + th[4] * y + th[5];
    const float ix = ((gx + 1.f) * W - 1.f) * 0.5f, iy = ((gy + 1.f) * H - 1.f) * 0.5f;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float ax = ix - fx0, ay = iy - fy0;
    const float w00 = (1.f - ax) * (1.f - ay), w01 = ax * (1.f - ay), w10 = (1.f - ax) * ay, w11 = ax * ay;
    const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
    const bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
    const float* s0 = src + (src_batch == 1 ? 0 : (long)n * C * H * W);
    for (int c = 0; c < C; ++c) {
        const float* sp = s0 + (long)c * H * W;
        float v = 0.f;
        if (vy0 && vx0) v += w00 * __ldg(sp + (long)y0 * W + x0);
        if (vy0 && vx1) v += w01 * __ldg(sp + (long)y0 * W + x0 + 1);
        if (vy1 && vx0) v += w10 * __ldg(sp + (long)(y0 + 1) * W + x0);
        if (vy1 && vx1) v += w11 * __ldg(sp + (long)(y0 + 1) * W + x0 + 1);
        dst[((long)n * C + c) * H * W + (long)i * W + j] = v;
    }
}
void k_affine_resample(const float* src, int src_batch, const float* theta, float* dst, int b, int C, int H, int W,
                       cudaStream_t st) {
    affine_resample_kernel<<<cdiv((long)b * H * W, 256), 256, 0, st>>>(src, src_batch, theta, dst, b, C, H, W); count_launch();
}

}  // namespace p2l
