// Depth-wise causal conv1d (+ bias, + SiLU) of the Mamba mixers, forward and backward, as sm_100a kernels.
//
// Reference semantics: `self.act(self.conv1d(x)[..., :seqlen])` with `nn.Conv1d(d, d, k, groups=d, padding=k-1)`
// (CXPMRG_Bench_MambaXray_VL/arm/Finetuning/mamba_simple.py:112-120, 673), which is what the un-vendored
// `causal_conv1d.causal_conv1d_fn(x, weight, bias, activation)` computes (:676-681):
//     pre[b, d, t] = bias[d] + sum_{k < K} w[d, k] * x[b, d, t - (K - 1) + k]        y = pre * sigmoid(pre)  (silu)
// HBM-bound (2 accesses per element forward, 3 backward).  Every thread owns 4 consecutive tokens of one (batch,
// channel) row (8- or 16-byte accesses, 256-512 B per warp instruction); the K - 1 halo tokens come from the
// neighbouring lane by shuffle.
// Backward = one CTA per channel: dx for all the batch rows of the channel, and dweight[d, :], dbias[d] reduced in
// registers -> shared memory -> one write: deterministic, no atomics.
#include <cuda_runtime.h>

#include <cstdio>
#include <type_traits>

#include "../../include/mia_selective_scan.h"
#include "scan_common.cuh"

namespace {

constexpr int kConvThreads = 256;
constexpr int kMaxW = 4;

struct ConvArgs {
    const void *x, *dy;
    void *y, *dx;
    const float *w, *bias;      // (dim, width) / (dim) fp32, bias may be null
    float *dw, *dbias;
    int batch, dim, L, width, silu;
    long long x_bs, x_ds, y_bs, y_ds, dy_bs, dy_ds, dx_bs, dx_ds;   // element strides; the sequence stride is 1
};

template <typename T> struct Tok4;   // 4 consecutive tokens <-> float[4]
template <> struct Tok4<float> {
    static __device__ __forceinline__ void ld(const void *p, float (&f)[4]) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
    static __device__ __forceinline__ void st(void *p, const float (&f)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Tok4<__nv_bfloat16> {
    static __device__ __forceinline__ void ld(const void *p, float (&f)[4]) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
        f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void st(void *p, const float (&f)[4]) {
        __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
        *reinterpret_cast<uint2 *>(p) = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
    }
};
template <> struct Tok4<__half> {
    static __device__ __forceinline__ void ld(const void *p, float (&f)[4]) {
        uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        const float2 a = __half22float2(*reinterpret_cast<__half2 *>(&v.x)), b = __half22float2(*reinterpret_cast<__half2 *>(&v.y));
        f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
    }
    static __device__ __forceinline__ void st(void *p, const float (&f)[4]) {
        __half2 a = __floats2half2_rn(f[0], f[1]), b = __floats2half2_rn(f[2], f[3]);
        *reinterpret_cast<uint2 *>(p) = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
    }
};

__device__ __forceinline__ float sigmoid_f(float v) { return mia::rcpf(1.f + mia::ex2f(-v * mia::kLog2e)); }

constexpr int kConvUnroll = 4;   // quads in flight per thread: one 8-byte load per thread cannot cover the HBM latency

// ------------------------------------------------------------------------------------------------ forward
// Rows (channels of one batch element) are contiguous, so a block's range of `rows_per_block` channels is one flat run
// of quads: quad i lives at base + 4 i elements, and only the channel (taps, from shared memory) and the row-start test
// need i / quads_per_row (a multiply-high).  grid = (channel ranges, batch).  The three halo tokens come from the
// neighbouring lane (shuffle) except at lane 0.
template <typename T, int kW, bool kSilu>
__global__ void __launch_bounds__(kConvThreads) causal_conv1d_fwd_kernel(const ConvArgs a, const int rows_per_block, const uint32_t qpr_magic) {
    constexpr int es = (int)sizeof(T);
    constexpr int U = kConvUnroll;
    extern __shared__ float taps[];                            // [rows_per_block][kW + 1]: w[0..kW-1], bias
    const int qpr = a.L >> 2;                                  // quads per row
    const int c0 = blockIdx.x * rows_per_block, nrows = min(rows_per_block, a.dim - c0);
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < nrows * (kW + 1); i += kConvThreads) {
        const int r = i / (kW + 1), k = i - r * (kW + 1);
        taps[i] = k < kW ? __ldg(a.w + (size_t)(c0 + r) * kW + k) : (a.bias ? __ldg(a.bias + c0 + r) : 0.f);
    }
    __syncthreads();
    const int nq = nrows * qpr;
    const char *xb = (const char *)a.x + ((size_t)b * a.x_bs + (size_t)c0 * a.L) * es;
    char *yb = (char *)a.y + ((size_t)b * a.y_bs + (size_t)c0 * a.L) * es;
    const int lane = threadIdx.x & 31;
    for (int i0 = 0; i0 < nq; i0 += U * kConvThreads) {
        float cur[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {                          // all the loads of the pass are issued before any use
            const int i = min(i0 + u * kConvThreads + (int)threadIdx.x, nq - 1);
            Tok4<T>::ld(xb + (size_t)i * (4 * es), cur[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int iu = i0 + u * kConvThreads + (int)threadIdx.x;
            const int i = min(iu, nq - 1);
            const int r = qpr_magic ? (int)__umulhi((uint32_t)i, qpr_magic) : i;   // i / qpr (magic 0: qpr == 1)
            const bool row_start = i == r * qpr;
            const float *tp = taps + r * (kW + 1);
            float prev[4], xv[8];
#pragma unroll
            for (int j = 1; j < 4; ++j) prev[j] = __shfl_up_sync(0xffffffffu, cur[u][j], 1);   // lane - 1 holds the previous quad ...
            if (lane == 0 && !row_start) Tok4<T>::ld(xb + (size_t)i * (4 * es) - 4 * es, prev);  // ... except for lane 0
            if (row_start) prev[1] = prev[2] = prev[3] = 0.f;                                  // causal zero padding
#pragma unroll
            for (int j = 1; j < 4; ++j) xv[j] = prev[j];
#pragma unroll
            for (int j = 0; j < 4; ++j) xv[4 + j] = cur[u][j];
            float y[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float acc = tp[kW];
#pragma unroll
                for (int k = 0; k < kW; ++k) acc = fmaf(tp[k], xv[j + 1 + k + (kMaxW - kW)], acc);
                y[j] = kSilu ? acc * sigmoid_f(acc) : acc;
            }
            if (iu < nq) Tok4<T>::st(yb + (size_t)i * (4 * es), y);
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward
// One CTA per channel.  The quads (batch b, quad q) of the channel are numbered j = b * qpr + q; a warp pass covers 31
// of them plus one look-ahead lane: lane l computes d pre of quad j = 31 k + l and hands its first three values to lane
// l - 1 (dx[t] needs d pre up to t + 3), so nothing is computed twice except by lane 31.
template <typename T, int kW, bool kSilu>
__global__ void __launch_bounds__(kConvThreads) causal_conv1d_bwd_kernel(const ConvArgs a, const uint32_t qpr_magic) {
    constexpr int es = (int)sizeof(T);
    constexpr int kWarps = kConvThreads / 32;
    constexpr int U = kConvUnroll;
    __shared__ float red[kWarps][kMaxW + 1];
    const int d = blockIdx.x;
    const int qpr = a.L >> 2;
    const int total = a.batch * qpr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float bias = a.bias ? __ldg(a.bias + d) : 0.f;
    float w[kW], dw[kW], db = 0.f;
#pragma unroll
    for (int k = 0; k < kW; ++k) { w[k] = __ldg(a.w + (size_t)d * kW + k); dw[k] = 0.f; }
    const char *xd = (const char *)a.x + (size_t)d * a.x_ds * es;
    const char *dyd = (const char *)a.dy + (size_t)d * a.dy_ds * es;
    char *dxd = (char *)a.dx + (size_t)d * a.dx_ds * es;
    for (int j00 = warp * 31; j00 < total; j00 += U * kWarps * 31) {
        int q[U], b[U];
        float cur[U][4], dyv[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {                          // all the loads of the pass are issued before any use
            const int j = min(j00 + u * kWarps * 31 + lane, total - 1);
            b[u] = qpr_magic ? (int)__umulhi((uint32_t)j, qpr_magic) : j;   // j / qpr (magic 0: qpr == 1)
            q[u] = j - b[u] * qpr;
            Tok4<T>::ld(xd + ((size_t)b[u] * a.x_bs + 4 * q[u]) * es, cur[u]);
            Tok4<T>::ld(dyd + ((size_t)b[u] * a.dy_bs + 4 * q[u]) * es, dyv[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (j00 + u * kWarps * 31 >= total) break;        // warp-uniform
            const bool valid = j00 + u * kWarps * 31 + lane < total;
            float prev[4], xv[8];
#pragma unroll
            for (int t = 1; t < 4; ++t) prev[t] = __shfl_up_sync(0xffffffffu, cur[u][t], 1);
            if (lane == 0 && q[u] > 0) Tok4<T>::ld(xd + ((size_t)b[u] * a.x_bs + 4 * q[u] - 4) * es, prev);
            if (q[u] == 0) prev[1] = prev[2] = prev[3] = 0.f;
#pragma unroll
            for (int t = 1; t < 4; ++t) xv[t] = prev[t];
#pragma unroll
            for (int t = 0; t < 4; ++t) xv[4 + t] = cur[u][t];
            float dp[7];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float g = valid ? dyv[u][t] : 0.f;
                if (kSilu) {
                    float pre = bias;
#pragma unroll
                    for (int k = 0; k < kW; ++k) pre = fmaf(w[k], xv[t + 1 + k + (kMaxW - kW)], pre);
                    const float s = sigmoid_f(pre);
                    g *= s * fmaf(pre, 1.f - s, 1.f);
                }
                dp[t] = g;
            }
            // d pre of the next quad of the same row (zero past the row end)
            const bool has_next = valid && q[u] + 1 < qpr;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const float nx = __shfl_down_sync(0xffffffffu, dp[t], 1);
                dp[4 + t] = has_next ? nx : 0.f;
            }
            if (valid && lane < 31) {                          // lane 31 only looked ahead for lane 30
                float dxv[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float acc = 0.f;                           // dx[t] = sum_k w[k] dpre[t + (kW - 1) - k]
#pragma unroll
                    for (int k = 0; k < kW; ++k) acc = fmaf(w[k], dp[t + (kW - 1) - k], acc);
                    dxv[t] = acc;
                    db += dp[t];
#pragma unroll
                    for (int k = 0; k < kW; ++k) dw[k] = fmaf(dp[t], xv[t + 1 + k + (kMaxW - kW)], dw[k]);
                }
                Tok4<T>::st(dxd + ((size_t)b[u] * a.dx_bs + 4 * q[u]) * es, dxv);
            }
        }
    }
    // block reduction in a fixed order: lanes (shuffle tree), then warps
    float vals[kW + 1];
#pragma unroll
    for (int k = 0; k < kW; ++k) vals[k] = dw[k];
    vals[kW] = db;
#pragma unroll
    for (int v = 0; v < kW + 1; ++v) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) vals[v] += __shfl_xor_sync(0xffffffffu, vals[v], off);
    }
    if (lane == 0) {
#pragma unroll
        for (int v = 0; v < kW + 1; ++v) red[warp][v] = vals[v];
    }
    __syncthreads();
    if (threadIdx.x < kW + 1) {
        float s = 0.f;
        for (int wv = 0; wv < kWarps; ++wv) s += red[wv][threadIdx.x];
        if (threadIdx.x == kW) {
            if (a.dbias) a.dbias[d] = s;
        } else {
            a.dw[(size_t)d * kW + threadIdx.x] = s;
        }
    }
}

// exact floor(i / d) for 0 <= i with i * d < 2^32 by multiply-high
uint32_t div_magic(int d) { return d == 1 ? 0u : (uint32_t)((0x100000000ULL + (uint64_t)d - 1) / (uint64_t)d); }

thread_local char g_conv_err[256] = "";

int conv_check(ConvArgs &a, const void *x, int batch, int dim, int L, int width, int dtype, const long long *strides, int nstrides) {
    if (!x || batch <= 0 || dim <= 0 || L <= 0) { snprintf(g_conv_err, sizeof(g_conv_err), "causal_conv1d: empty or null input"); return MIA_EINVAL; }
    if (width < 1 || width > kMaxW) { snprintf(g_conv_err, sizeof(g_conv_err), "causal_conv1d only supports width between 1 and 4"); return MIA_EINVAL; }
    if (dtype != MIA_F32 && dtype != MIA_F16 && dtype != MIA_BF16) { snprintf(g_conv_err, sizeof(g_conv_err), "causal_conv1d: bad dtype"); return MIA_EINVAL; }
    if (L % 4) { snprintf(g_conv_err, sizeof(g_conv_err), "causal_conv1d: seqlen must be a multiple of 4 (got %d)", L); return MIA_EINVAL; }
    for (int i = 0; i < nstrides; ++i)
        if (strides[i] % 4) { snprintf(g_conv_err, sizeof(g_conv_err), "causal_conv1d: batch / channel strides must be multiples of 4 elements"); return MIA_EINVAL; }
    (void)a;
    return MIA_OK;
}

template <typename F>
int conv_dispatch(int dtype, F &&f) {
    switch (dtype) {
        case MIA_F32: return f((float *)nullptr);
        case MIA_F16: return f((__half *)nullptr);
        default: return f((__nv_bfloat16 *)nullptr);
    }
}

template <typename T>
void launch_conv_fwd(const ConvArgs &a, dim3 grid, int rpb, uint32_t magic, size_t smem, cudaStream_t st) {
#define MIA_CONV_FWD(W)                                                                              \
    do {                                                                                             \
        if (a.silu) causal_conv1d_fwd_kernel<T, W, true><<<grid, kConvThreads, smem, st>>>(a, rpb, magic);  \
        else causal_conv1d_fwd_kernel<T, W, false><<<grid, kConvThreads, smem, st>>>(a, rpb, magic);        \
    } while (0)
    switch (a.width) {
        case 4: MIA_CONV_FWD(4); break;
        case 3: MIA_CONV_FWD(3); break;
        case 2: MIA_CONV_FWD(2); break;
        default: MIA_CONV_FWD(1); break;
    }
#undef MIA_CONV_FWD
}

template <typename T>
void launch_conv_bwd(const ConvArgs &a, uint32_t magic, cudaStream_t st) {
#define MIA_CONV_BWD(W)                                                                              \
    do {                                                                                             \
        if (a.silu) causal_conv1d_bwd_kernel<T, W, true><<<a.dim, kConvThreads, 0, st>>>(a, magic);  \
        else causal_conv1d_bwd_kernel<T, W, false><<<a.dim, kConvThreads, 0, st>>>(a, magic);        \
    } while (0)
    switch (a.width) {
        case 4: MIA_CONV_BWD(4); break;
        case 3: MIA_CONV_BWD(3); break;
        case 2: MIA_CONV_BWD(2); break;
        default: MIA_CONV_BWD(1); break;
    }
#undef MIA_CONV_BWD
}

}  // namespace

extern "C" {

const char *mia_conv_last_error(void) { return g_conv_err; }

int mia_causal_conv1d_fwd(const void *x, const float *weight, const float *bias, void *y, int batch, int dim, int seqlen, int width,
                          int silu, int dtype, long long x_batch_stride, long long x_d_stride, long long y_batch_stride,
                          long long y_d_stride, void *cuda_stream) {
    ConvArgs a{};
    const long long st[4] = {x_batch_stride, x_d_stride, y_batch_stride, y_d_stride};
    if (int rc = conv_check(a, x, batch, dim, seqlen, width, dtype, st, 4)) return rc;
    const int es = dtype == MIA_F32 ? 4 : 2;
    if (!weight || !y || (((uintptr_t)x | (uintptr_t)y) & (4 * es - 1))) {
        snprintf(g_conv_err, sizeof(g_conv_err), "causal_conv1d_fwd: null weight / y or x / y not aligned to 4 elements");
        return MIA_EINVAL;
    }
    a.x = x; a.y = y; a.w = weight; a.bias = bias; a.batch = batch; a.dim = dim; a.L = seqlen; a.width = width; a.silu = silu;
    a.x_bs = x_batch_stride; a.x_ds = x_d_stride; a.y_bs = y_batch_stride; a.y_ds = y_d_stride;
    if (x_d_stride != seqlen || y_d_stride != seqlen) {
        snprintf(g_conv_err, sizeof(g_conv_err), "causal_conv1d_fwd: the channel rows of x and y must be contiguous (stride == seqlen)");
        return MIA_EINVAL;
    }
    // rows per block: ~2 passes of kConvUnroll quads per thread
    const int qpr = seqlen / 4;
    int rpb = (2 * kConvUnroll * kConvThreads + qpr - 1) / qpr;
    if (rpb > dim) rpb = dim;
    if ((long long)rpb * qpr * qpr >= (1LL << 32) || batch > 65535) {
        snprintf(g_conv_err, sizeof(g_conv_err), "causal_conv1d: sequence or batch too large");
        return MIA_EINVAL;
    }
    const dim3 grid((dim + rpb - 1) / rpb, batch);
    const size_t smem = (size_t)rpb * (width + 1) * sizeof(float);
    const int rc = conv_dispatch(dtype, [&](auto *tag) {
        using T = typename std::remove_pointer<decltype(tag)>::type;
        launch_conv_fwd<T>(a, grid, rpb, div_magic(qpr), smem, (cudaStream_t)cuda_stream);
        return (int)cudaGetLastError();
    });
    if (rc != 0) { snprintf(g_conv_err, sizeof(g_conv_err), "causal_conv1d_fwd launch: %s", cudaGetErrorString((cudaError_t)rc)); return MIA_ECUDA; }
    return MIA_OK;
}

int mia_causal_conv1d_bwd(const void *x, const float *weight, const float *bias, const void *dy, void *dx, float *dweight, float *dbias,
                          int batch, int dim, int seqlen, int width, int silu, int dtype, long long x_batch_stride, long long x_d_stride,
                          long long dy_batch_stride, long long dy_d_stride, long long dx_batch_stride, long long dx_d_stride,
                          void *cuda_stream) {
    ConvArgs a{};
    const long long st[6] = {x_batch_stride, x_d_stride, dy_batch_stride, dy_d_stride, dx_batch_stride, dx_d_stride};
    if (int rc = conv_check(a, x, batch, dim, seqlen, width, dtype, st, 6)) return rc;
    const int es = dtype == MIA_F32 ? 4 : 2;
    if (!weight || !dy || !dx || !dweight || (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) & (4 * es - 1))) {
        snprintf(g_conv_err, sizeof(g_conv_err), "causal_conv1d_bwd: null pointer or x / dy / dx not aligned to 4 elements");
        return MIA_EINVAL;
    }
    a.x = x; a.dy = dy; a.dx = dx; a.w = weight; a.bias = bias; a.dw = dweight; a.dbias = dbias;
    a.batch = batch; a.dim = dim; a.L = seqlen; a.width = width; a.silu = silu;
    a.x_bs = x_batch_stride; a.x_ds = x_d_stride; a.dy_bs = dy_batch_stride; a.dy_ds = dy_d_stride; a.dx_bs = dx_batch_stride; a.dx_ds = dx_d_stride;
    const int qpr = seqlen / 4;
    if ((long long)batch * qpr * qpr >= (1LL << 32)) {
        snprintf(g_conv_err, sizeof(g_conv_err), "causal_conv1d: batch * sequence too large");
        return MIA_EINVAL;
    }
    const int rc = conv_dispatch(dtype, [&](auto *tag) {
        using T = typename std::remove_pointer<decltype(tag)>::type;
        launch_conv_bwd<T>(a, div_magic(qpr), (cudaStream_t)cuda_stream);
        return (int)cudaGetLastError();
    });
    if (rc != 0) { snprintf(g_conv_err, sizeof(g_conv_err), "causal_conv1d_bwd launch: %s", cudaGetErrorString((cudaError_t)rc)); return MIA_ECUDA; }
    return MIA_OK;
}

}  // extern "C"
