// Depth-wise 3x3 conv2d (+ bias, + SiLU) of SS2D, forward and backward, as sm_100a kernels.
//
// Reference: `x = self.act(self.conv2d(x))` with `nn.Conv2d(d_inner, d_inner, groups=d_inner, kernel_size=3, padding=1)`
// on channel-first (B, C, H, W) activations (R2GenCSR/VMamba/classification/models/vmamba.py:574-582, 1120-1122).
//     pre[b, c, h, w] = bias[c] + sum_{i, j in 0..2} wgt[c, i, j] * x[b, c, h + i - 1, w + j - 1]     y = pre * sigmoid(pre)
// HBM-bound (2 tensor passes forward, 3 backward).  A (batch, channel) plane is contiguous: planes are staged whole in
// shared memory with coalesced copies and the 3x3 stencil reads them from there (border taps predicated to zero).
// Backward = one CTA per channel over all the batch planes of the channel: d pre goes through shared memory for the
// transposed stencil (dx), dweight[c, :, :] and dbias[c] are reduced in registers -> shared memory -> one write:
// deterministic, no atomics.
#include <cuda_runtime.h>

#include <cstdio>
#include <type_traits>

#include "../../include/mia_selective_scan.h"
#include "scan_common.cuh"

namespace {

constexpr int kDwThreads = 256;

struct DwArgs {
    const void *x, *dy;
    void *y, *dx;
    const float *w, *bias;      // (C, 9) / (C) fp32, bias may be null
    float *dw, *dbias;
    int batch, C, H, W, silu;
    int planes_per_block;       // forward: planes staged per block pass; backward: batch planes per pass
    uint32_t magic_hw, magic_w; // multiply-high reciprocals of H*W and W (0: divisor 1)
};

__device__ __forceinline__ int fast_div(int i, uint32_t magic) { return magic ? (int)__umulhi((uint32_t)i, magic) : i; }
__device__ __forceinline__ float dw_sigmoid(float v) { return mia::rcpf(1.f + mia::ex2f(-v * mia::kLog2e)); }

// 3x3 stencil around (h, w) of one plane in shared memory (raw element type), zero outside the plane
template <typename T>
__device__ __forceinline__ float stencil(const typename mia::Cvt<T>::raw *pl, int h, int w, int H, int W, const float (&k)[9], float acc) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int hh = h + i - 1;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int ww = w + j - 1;
            const bool in = (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W;
            const float v = in ? mia::Cvt<T>::to_f(pl[hh * W + ww]) : 0.f;
            acc = fmaf(k[i * 3 + j], v, acc);
        }
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------ forward
template <typename T, bool kSilu>
__global__ void __launch_bounds__(kDwThreads) dwconv2d_fwd_kernel(const DwArgs a) {
    using raw = typename mia::Cvt<T>::raw;
    extern __shared__ __align__(16) char dsm[];
    const int HW = a.H * a.W, P = a.planes_per_block;
    raw *sx = reinterpret_cast<raw *>(dsm);                                   // [P][HW]
    float *taps = reinterpret_cast<float *>(dsm + (((size_t)P * HW * sizeof(raw) + 15) & ~(size_t)15));   // [P][10]
    const int n_planes = a.batch * a.C;
    for (int p0 = blockIdx.x * P; p0 < n_planes; p0 += gridDim.x * P) {
        const int np = min(P, n_planes - p0);
        const raw *gx = reinterpret_cast<const raw *>(a.x) + (size_t)p0 * HW;
        for (int i = threadIdx.x; i < np * HW; i += kDwThreads) sx[i] = gx[i];
        for (int i = threadIdx.x; i < np * 10; i += kDwThreads) {
            const int pl = i / 10, k = i - pl * 10, c = (p0 + pl) % a.C;
            taps[i] = k < 9 ? __ldg(a.w + (size_t)c * 9 + k) : (a.bias ? __ldg(a.bias + c) : 0.f);
        }
        __syncthreads();
        raw *gy = reinterpret_cast<raw *>(a.y) + (size_t)p0 * HW;
        for (int i = threadIdx.x; i < np * HW; i += kDwThreads) {
            const int pl = fast_div(i, a.magic_hw), l = i - pl * HW;
            const int h = fast_div(l, a.magic_w), w = l - h * a.W;
            float k[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) k[t] = taps[pl * 10 + t];
            const float pre = stencil<T>(sx + pl * HW, h, w, a.H, a.W, k, taps[pl * 10 + 9]);
            gy[i] = mia::Cvt<T>::from_f(kSilu ? pre * dw_sigmoid(pre) : pre);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ backward
template <typename T, bool kSilu>
__global__ void __launch_bounds__(kDwThreads) dwconv2d_bwd_kernel(const DwArgs a) {
    using raw = typename mia::Cvt<T>::raw;
    extern __shared__ __align__(16) char dsm[];
    __shared__ float red[kDwThreads / 32][10];
    const int HW = a.H * a.W, P = a.planes_per_block;
    raw *sx = reinterpret_cast<raw *>(dsm);                                   // [P][HW]
    float *sdp = reinterpret_cast<float *>(dsm + (((size_t)P * HW * sizeof(raw) + 15) & ~(size_t)15));    // [P][HW] d pre
    const int c = blockIdx.x;
    float k[9], acc[10];
#pragma unroll
    for (int t = 0; t < 9; ++t) { k[t] = __ldg(a.w + (size_t)c * 9 + t); acc[t] = 0.f; }
    acc[9] = 0.f;
    const float bias = a.bias ? __ldg(a.bias + c) : 0.f;
    for (int b0 = 0; b0 < a.batch; b0 += P) {
        const int np = min(P, a.batch - b0);
        for (int i = threadIdx.x; i < np * HW; i += kDwThreads) {
            const int pl = fast_div(i, a.magic_hw), l = i - pl * HW;
            sx[i] = reinterpret_cast<const raw *>(a.x)[((size_t)(b0 + pl) * a.C + c) * HW + l];
        }
        __syncthreads();
        // d pre = dy * silu'(pre); dweight / dbias partial sums
        for (int i = threadIdx.x; i < np * HW; i += kDwThreads) {
            const int pl = fast_div(i, a.magic_hw), l = i - pl * HW;
            const int h = fast_div(l, a.magic_w), w = l - h * a.W;
            float g = mia::Cvt<T>::to_f(reinterpret_cast<const raw *>(a.dy)[((size_t)(b0 + pl) * a.C + c) * HW + l]);
            const raw *plx = sx + pl * HW;
            if (kSilu) {
                const float pre = stencil<T>(plx, h, w, a.H, a.W, k, bias);
                const float s = dw_sigmoid(pre);
                g *= s * fmaf(pre, 1.f - s, 1.f);
            }
            sdp[i] = g;
            acc[9] += g;
#pragma unroll
            for (int ii = 0; ii < 3; ++ii) {
                const int hh = h + ii - 1;
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    const int ww = w + jj - 1;
                    const bool in = (unsigned)hh < (unsigned)a.H && (unsigned)ww < (unsigned)a.W;
                    acc[ii * 3 + jj] = fmaf(g, in ? mia::Cvt<T>::to_f(plx[hh * a.W + ww]) : 0.f, acc[ii * 3 + jj]);
                }
            }
        }
        __syncthreads();
        // dx[h, w] = sum_{i, j} wgt[i, j] * d pre[h - i + 1, w - j + 1]   (transposed stencil)
        for (int i = threadIdx.x; i < np * HW; i += kDwThreads) {
            const int pl = fast_div(i, a.magic_hw), l = i - pl * HW;
            const int h = fast_div(l, a.magic_w), w = l - h * a.W;
            const float *pld = sdp + pl * HW;
            float v = 0.f;
#pragma unroll
            for (int ii = 0; ii < 3; ++ii) {
                const int hh = h - ii + 1;
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    const int ww = w - jj + 1;
                    const bool in = (unsigned)hh < (unsigned)a.H && (unsigned)ww < (unsigned)a.W;
                    v = fmaf(k[ii * 3 + jj], in ? pld[hh * a.W + ww] : 0.f, v);
                }
            }
            reinterpret_cast<raw *>(a.dx)[((size_t)(b0 + pl) * a.C + c) * HW + l] = mia::Cvt<T>::from_f(v);
        }
        __syncthreads();
    }
    // block reduction in a fixed order: lanes (shuffle tree), then warps
#pragma unroll
    for (int v = 0; v < 10; ++v) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], off);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int v = 0; v < 10; ++v) red[warp][v] = acc[v];
    }
    __syncthreads();
    if (threadIdx.x < 10) {
        float s = 0.f;
        for (int wv = 0; wv < kDwThreads / 32; ++wv) s += red[wv][threadIdx.x];
        if (threadIdx.x == 9) {
            if (a.dbias) a.dbias[c] = s;
        } else {
            a.dw[(size_t)c * 9 + threadIdx.x] = s;
        }
    }
}

thread_local char g_dw_err[256] = "";
uint32_t dw_magic(int d) { return d == 1 ? 0u : (uint32_t)((0x100000000ULL + (uint64_t)d - 1) / (uint64_t)d); }

int dw_check(const void *x, int batch, int C, int H, int W, int dtype) {
    if (!x || batch <= 0 || C <= 0 || H <= 0 || W <= 0) { snprintf(g_dw_err, sizeof(g_dw_err), "dwconv2d: empty or null input"); return MIA_EINVAL; }
    if (dtype != MIA_F32 && dtype != MIA_F16 && dtype != MIA_BF16) { snprintf(g_dw_err, sizeof(g_dw_err), "dwconv2d: bad dtype"); return MIA_EINVAL; }
    if ((long long)H * W > 16384) { snprintf(g_dw_err, sizeof(g_dw_err), "dwconv2d: planes larger than 16384 elements are not supported"); return MIA_EINVAL; }
    if ((long long)batch * C >= (1LL << 31) / 2) { snprintf(g_dw_err, sizeof(g_dw_err), "dwconv2d: batch*channels too large"); return MIA_EINVAL; }
    return MIA_OK;
}

template <typename F>
int dw_dispatch(int dtype, F &&f) {
    switch (dtype) {
        case MIA_F32: return f((float *)nullptr);
        case MIA_F16: return f((__half *)nullptr);
        default: return f((__nv_bfloat16 *)nullptr);
    }
}

}  // namespace

extern "C" {

const char *mia_dwconv2d_last_error(void) { return g_dw_err; }

int mia_dwconv2d_fwd(const void *x, const float *weight, const float *bias, void *y, int batch, int channels, int H, int W, int silu,
                     int dtype, void *cuda_stream) {
    if (int rc = dw_check(x, batch, channels, H, W, dtype)) return rc;
    if (!weight || !y) { snprintf(g_dw_err, sizeof(g_dw_err), "dwconv2d_fwd: null weight or y"); return MIA_EINVAL; }
    const int es = dtype == MIA_F32 ? 4 : 2, HW = H * W;
    DwArgs a{};
    a.x = x; a.y = y; a.w = weight; a.bias = bias; a.batch = batch; a.C = channels; a.H = H; a.W = W; a.silu = silu;
    int P = 2048 / HW;
    if (P < 1) P = 1;
    if (P > 16) P = 16;
    a.planes_per_block = P;
    a.magic_hw = dw_magic(HW); a.magic_w = dw_magic(W);
    const size_t smem = (((size_t)P * HW * es + 15) & ~(size_t)15) + (size_t)P * 10 * sizeof(float);
    const long long n_planes = (long long)batch * channels;
    long long blocks = (n_planes + P - 1) / P;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    const int rc = dw_dispatch(dtype, [&](auto *tag) {
        using T = typename std::remove_pointer<decltype(tag)>::type;
        auto kern = silu ? &dwconv2d_fwd_kernel<T, true> : &dwconv2d_fwd_kernel<T, false>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<(int)blocks, kDwThreads, smem, (cudaStream_t)cuda_stream>>>(a);
        return (int)cudaGetLastError();
    });
    if (rc != 0) { snprintf(g_dw_err, sizeof(g_dw_err), "dwconv2d_fwd launch: %s", cudaGetErrorString((cudaError_t)rc)); return MIA_ECUDA; }
    return MIA_OK;
}

int mia_dwconv2d_bwd(const void *x, const float *weight, const float *bias, const void *dy, void *dx, float *dweight, float *dbias,
                     int batch, int channels, int H, int W, int silu, int dtype, void *cuda_stream) {
    if (int rc = dw_check(x, batch, channels, H, W, dtype)) return rc;
    if (!weight || !dy || !dx || !dweight) { snprintf(g_dw_err, sizeof(g_dw_err), "dwconv2d_bwd: null pointer"); return MIA_EINVAL; }
    const int es = dtype == MIA_F32 ? 4 : 2, HW = H * W;
    DwArgs a{};
    a.x = x; a.dy = dy; a.dx = dx; a.w = weight; a.bias = bias; a.dw = dweight; a.dbias = dbias;
    a.batch = batch; a.C = channels; a.H = H; a.W = W; a.silu = silu;
    int P = 2048 / HW;
    if (P < 1) P = 1;
    if (P > batch) P = batch;
    a.planes_per_block = P;
    a.magic_hw = dw_magic(HW); a.magic_w = dw_magic(W);
    const size_t smem = (((size_t)P * HW * es + 15) & ~(size_t)15) + (size_t)P * HW * sizeof(float);
    const int rc = dw_dispatch(dtype, [&](auto *tag) {
        using T = typename std::remove_pointer<decltype(tag)>::type;
        auto kern = silu ? &dwconv2d_bwd_kernel<T, true> : &dwconv2d_bwd_kernel<T, false>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<channels, kDwThreads, smem, (cudaStream_t)cuda_stream>>>(a);
        return (int)cudaGetLastError();
    });
    if (rc != 0) { snprintf(g_dw_err, sizeof(g_dw_err), "dwconv2d_bwd launch: %s", cudaGetErrorString((cudaError_t)rc)); return MIA_ECUDA; }
    return MIA_OK;
}

}  // extern "C"
