// CrossScan / CrossMerge of VMamba's SS2D as single-pass sm_100a kernels.
//
// Reference: R2GenCSR/VMamba/classification/models/vmamba.py:25-67 (torch: 3-5 full-tensor copies each way) and
// csm_triton.py:7-235 (Triton; the north star forbids Triton).  One CTA stages a handful of (batch, channel) planes in
// shared memory and emits all four scan orders (row-major, column-major and their reversals) with coalesced global
// accesses; the merge is the exact adjoint (accumulated in fp32, rounded once).  HBM-bound: 1 read + 4 writes (scan),
// 4 reads + 1 write (merge) per element.
#include <cuda_runtime.h>

#include "../../include/mia_selective_scan.h"
#include "scan_common.cuh"

namespace {

constexpr int kPlanes = 4;      // planes per CTA
constexpr int kCsThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kCsThreads) cross_scan_kernel(const typename mia::Cvt<T>::raw *__restrict__ x,
                                                                typename mia::Cvt<T>::raw *__restrict__ xs, int n_planes, int C,
                                                                int H, int W) {
    using raw = typename mia::Cvt<T>::raw;
    extern __shared__ __align__(16) char smem_raw[];
    raw *s = reinterpret_cast<raw *>(smem_raw);
    const int L = H * W;
    const int p0 = blockIdx.x * kPlanes, np = min(kPlanes, n_planes - p0);
    for (int i = threadIdx.x; i < np * L; i += kCsThreads) s[i] = x[(size_t)p0 * L + i];
    __syncthreads();
    for (int i = threadIdx.x; i < np * L; i += kCsThreads) {
        const int pl = i / L, l = i - pl * L;
        const int p = p0 + pl, b = p / C, c = p - b * C;
        const raw v = s[i];
        const raw vt = s[pl * L + (l % H) * W + l / H];      // column-major order: position l' = w * H + h holds x[h][w]
        raw *o = xs + ((size_t)b * 4 * C + c) * L;
        const size_t kstep = (size_t)C * L;
        o[l] = v;
        o[kstep + l] = vt;
        o[2 * kstep + (L - 1 - l)] = v;
        o[3 * kstep + (L - 1 - l)] = vt;
    }
}

template <typename T>
__global__ void __launch_bounds__(kCsThreads) cross_merge_kernel(const typename mia::Cvt<T>::raw *__restrict__ ys,
                                                                 typename mia::Cvt<T>::raw *__restrict__ y, int n_planes, int C,
                                                                 int H, int W) {
    using raw = typename mia::Cvt<T>::raw;
    extern __shared__ __align__(16) char smem_raw[];
    float *t = reinterpret_cast<float *>(smem_raw);          // column-major partial: ys1 + flip(ys3)
    const int L = H * W;
    const int p0 = blockIdx.x * kPlanes, np = min(kPlanes, n_planes - p0);
    const size_t kstep = (size_t)C * L;
    for (int i = threadIdx.x; i < np * L; i += kCsThreads) {
        const int pl = i / L, l = i - pl * L;
        const int p = p0 + pl, b = p / C, c = p - b * C;
        const raw *in = ys + ((size_t)b * 4 * C + c) * L;
        t[i] = mia::Cvt<T>::to_f(in[kstep + l]) + mia::Cvt<T>::to_f(in[3 * kstep + (L - 1 - l)]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < np * L; i += kCsThreads) {
        const int pl = i / L, l = i - pl * L;
        const int p = p0 + pl, b = p / C, c = p - b * C;
        const raw *in = ys + ((size_t)b * 4 * C + c) * L;
        const int h = l / W, w = l - h * W;
        const float v = mia::Cvt<T>::to_f(in[l]) + mia::Cvt<T>::to_f(in[2 * kstep + (L - 1 - l)]) + t[pl * L + w * H + h];
        y[(size_t)p * L + l] = mia::Cvt<T>::from_f(v);
    }
}

thread_local char g_cs_err[256] = "";

template <typename T>
int launch_cs(bool merge, const void *in, void *out, int B, int C, int H, int W, cudaStream_t stream) {
    using raw = typename mia::Cvt<T>::raw;
    const int n_planes = B * C, L = H * W;
    const size_t smem = (size_t)kPlanes * L * (merge ? sizeof(float) : sizeof(raw));
    if (smem > 200 * 1024) return MIA_EINVAL;
    const int grid = (n_planes + kPlanes - 1) / kPlanes;
    cudaError_t e;
    if (merge) {
        auto k = &cross_merge_kernel<T>;
        if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return MIA_ECUDA;
        k<<<grid, kCsThreads, smem, stream>>>(reinterpret_cast<const raw *>(in), reinterpret_cast<raw *>(out), n_planes, C, H, W);
    } else {
        auto k = &cross_scan_kernel<T>;
        if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return MIA_ECUDA;
        k<<<grid, kCsThreads, smem, stream>>>(reinterpret_cast<const raw *>(in), reinterpret_cast<raw *>(out), n_planes, C, H, W);
    }
    return cudaGetLastError() == cudaSuccess ? MIA_OK : MIA_ECUDA;
}

int cs_dispatch(bool merge, const void *in, void *out, int B, int C, int H, int W, int dtype, void *stream) {
    if (!in || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0) return MIA_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case MIA_F32: return launch_cs<float>(merge, in, out, B, C, H, W, st);
        case MIA_F16: return launch_cs<__half>(merge, in, out, B, C, H, W, st);
        case MIA_BF16: return launch_cs<__nv_bfloat16>(merge, in, out, B, C, H, W, st);
        default: return MIA_EINVAL;
    }
}

}  // namespace

extern "C" {
int mia_cross_scan(const void *x, void *xs, int batch, int channels, int H, int W, int dtype, void *cuda_stream) {
    return cs_dispatch(false, x, xs, batch, channels, H, W, dtype, cuda_stream);
}
int mia_cross_merge(const void *ys, void *y, int batch, int channels, int H, int W, int dtype, void *cuda_stream) {
    return cs_dispatch(true, ys, y, batch, channels, H, W, dtype, cuda_stream);
}
}
