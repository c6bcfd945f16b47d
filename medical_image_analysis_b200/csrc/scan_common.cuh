// Device-side building blocks shared by the forward and backward selective-scan kernels (sm_100a).
//
// Design: persistent CTAs; one producer warp stages (u, delta, [z, dout, out], B, C)
// tiles into shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) through a
// multi-stage full/empty ring; consumer warps own rows.  A row chunk is LPR lanes x 8 consecutive tokens:
// every lane scans its 8 tokens serially in registers, the lane aggregates are combined with a
// warp-shuffle scan of (a, b) pairs under the operator (a, b) o (a', b') = (a a', a' b + b')
// (the algebra of selective_scan_common.h:91-96 in the reference), chunks are chained through a carry.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mia {

constexpr int kTok = 8;  // tokens per lane per chunk
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kMaxStages = 8;

// ------------------------------------------------------------------------------------------------
// Kernel argument block (built on the host by scan_api.cu).
struct ScanArgs {
    // problem
    int batch, dim, L, N, G, delta_dim;
    int rows_per_group, delta_ratio;
    int softplus, has_z, out_f32;      // out_f32: dtype of out/out_z (fwd) or dout/out_saved (bwd) is float
    // tiling
    int RT, tiles_per_group, LPR, CH, n_chunks, n_items, stages, n_consumer_warps;
    // span-merge flags (whole tile is one contiguous span in global memory)
    int flat_u, flat_delta, flat_z, flat_dout, flat_osaved, flat_B, flat_C;
    // shared-memory layout (bytes)
    int row_pitch, rowo_pitch, bc_pitch;
    int off_u, off_delta, off_z, off_dout, off_osaved, off_B, off_C, stage_bytes;
    int off_bars, off_carry, off_red, smem_bytes;
    // pointers
    const void *u, *delta, *A, *B, *C, *D, *delta_bias, *z;
    void *out, *out_z;
    float *x;
    const void *dout, *out_saved;
    void *du, *ddelta, *dz;
    float *part_dA, *part_dD, *part_dbias;   // (batch, dim, N) / (batch, dim) / (batch, dim) f32 partials
    float *acc_dB, *acc_dC;                  // N <= 2: (batch*G*tiles, N, L) partials ; else (batch, G, N, L) atomics
    float *ddelta_full;                      // (batch, dim, L) f32 when delta_ratio > 1
    int bc_atomic;
    // strides (elements)
    long long u_bs, u_ds, delta_bs, delta_ds, z_bs, z_ds, A_ds, A_ns;
    long long B_bs, B_gs, B_ns, C_bs, C_gs, C_ns;
    long long out_bs, out_ds, outz_bs, outz_ds;
    long long dout_bs, dout_ds, osaved_bs, osaved_ds;
    long long du_bs, du_ds, dd_bs, dd_ds, dz_bs, dz_ds;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA bulk copy, fast math.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// 1-D TMA: global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// softplus with the reference's threshold (fwd_kernel_oflex.cuh:126: x <= 20 ? log1p(exp(x)) : x) plus, optionally, its
// derivative sigmoid(x) (bwd_kernel_oflex.cuh:252-257).  log1p is evaluated by a series for small e so that the usual
// Mamba range (dt ~ 1e-3..1e-1) keeps full relative precision without the libm log1pf.
template <bool kWithSigmoid>
__device__ __forceinline__ float softplus_f(float x, float &sig) {
    const float e = ex2f(x * kLog2e);
    const float s = 1.f + e;
    const float series = e * (1.f + e * (-0.5f + e * (0.33333334f + e * (-0.25f + e * 0.2f))));
    float sp = e < 0.06f ? series : lg2f(s) * kLn2;
    if (kWithSigmoid) sig = x <= 20.f ? e * rcpf(s) : 1.f;
    return x <= 20.f ? sp : x;
}

// ------------------------------------------------------------------------------------------------
// dtype traits + 8-token vector access with runtime alignment dispatch (the dispatch is warp-uniform
// for a row chunk because every lane sits a multiple of 16 (or 32) bytes from the row start).
template <typename T> struct Cvt;
template <> struct Cvt<float> {
    using raw = uint32_t;
    static __device__ __forceinline__ float to_f(raw v) { return __uint_as_float(v); }
    static __device__ __forceinline__ raw from_f(float v) { return __float_as_uint(v); }
};
template <> struct Cvt<__half> {
    using raw = uint16_t;
    static __device__ __forceinline__ float to_f(raw v) { return __half2float(__ushort_as_half(v)); }
    static __device__ __forceinline__ raw from_f(float v) { return __half_as_ushort(__float2half_rn(v)); }
};
template <> struct Cvt<__nv_bfloat16> {
    using raw = uint16_t;
    static __device__ __forceinline__ float to_f(raw v) { return __uint_as_float((uint32_t)v << 16); }
    static __device__ __forceinline__ raw from_f(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
};

template <typename T> struct Pack8 { typename Cvt<T>::raw v[kTok]; };

// load 8 consecutive T from (possibly only element-aligned) address p (shared or global) into floats
template <typename T>
__device__ __forceinline__ void ld8(const void *p, float (&f)[kTok]) {
    constexpr int kBytes = kTok * (int)sizeof(T);
    union { Pack8<T> t; uint4 q[kBytes / 16]; uint2 d[kBytes / 8]; uint32_t w[kBytes / 4]; } buf;
    const uintptr_t a = (uintptr_t)p;
    if ((a & 15) == 0) {
#pragma unroll
        for (int i = 0; i < kBytes / 16; ++i) buf.q[i] = reinterpret_cast<const uint4 *>(p)[i];
    } else if ((a & 7) == 0) {
#pragma unroll
        for (int i = 0; i < kBytes / 8; ++i) buf.d[i] = reinterpret_cast<const uint2 *>(p)[i];
    } else if ((a & 3) == 0) {
#pragma unroll
        for (int i = 0; i < kBytes / 4; ++i) buf.w[i] = reinterpret_cast<const uint32_t *>(p)[i];
    } else {
#pragma unroll
        for (int i = 0; i < kTok; ++i) buf.t.v[i] = reinterpret_cast<const typename Cvt<T>::raw *>(p)[i];
    }
#pragma unroll
    for (int i = 0; i < kTok; ++i) f[i] = Cvt<T>::to_f(buf.t.v[i]);
}

// store up to 8 consecutive T to global memory (nvalid may be < 8 at the sequence tail)
template <typename T>
__device__ __forceinline__ void st8(void *p, const float (&f)[kTok], int nvalid) {
    constexpr int kBytes = kTok * (int)sizeof(T);
    union { Pack8<T> t; uint4 q[kBytes / 16]; uint2 d[kBytes / 8]; uint32_t w[kBytes / 4]; } buf;
#pragma unroll
    for (int i = 0; i < kTok; ++i) buf.t.v[i] = Cvt<T>::from_f(f[i]);
    const uintptr_t a = (uintptr_t)p;
    if (nvalid >= kTok) {
        if ((a & 15) == 0) {
#pragma unroll
            for (int i = 0; i < kBytes / 16; ++i) reinterpret_cast<uint4 *>(p)[i] = buf.q[i];
            return;
        } else if ((a & 7) == 0) {
#pragma unroll
            for (int i = 0; i < kBytes / 8; ++i) reinterpret_cast<uint2 *>(p)[i] = buf.d[i];
            return;
        } else if ((a & 3) == 0) {
#pragma unroll
            for (int i = 0; i < kBytes / 4; ++i) reinterpret_cast<uint32_t *>(p)[i] = buf.w[i];
            return;
        }
    }
#pragma unroll
    for (int i = 0; i < kTok; ++i)
        if (i < nvalid) reinterpret_cast<typename Cvt<T>::raw *>(p)[i] = buf.t.v[i];
}

// ------------------------------------------------------------------------------------------------
// Producer side: copy `nbytes` from global `src` into the shared slot so that source byte k lands at
// slot + (src & 15) + k.  The 16-byte-aligned interior goes through one TMA bulk copy; the (< 16 B) head and
// tail are copied element-wise by the calling lane.  Returns the bytes the bulk copy will complete on `bar`.
__device__ __forceinline__ uint32_t stage_span(char *slot, const char *src, uint32_t nbytes, uint64_t *bar, int esize) {
    const uint32_t mis = (uint32_t)((uintptr_t)src & 15u);
    char *dst = slot + mis;
    uint32_t head = mis ? (16u - mis) : 0u;
    if (head > nbytes) head = nbytes;
    const uint32_t body = (nbytes - head) & ~15u;
    const uint32_t tail = nbytes - head - body;
    if (body) bulk_g2s(dst + head, src + head, body, bar);
    if (esize == 2) {
        for (uint32_t k = 0; k < head; k += 2) *reinterpret_cast<uint16_t *>(dst + k) = *reinterpret_cast<const uint16_t *>(src + k);
        for (uint32_t k = head + body; k < head + body + tail; k += 2)
            *reinterpret_cast<uint16_t *>(dst + k) = *reinterpret_cast<const uint16_t *>(src + k);
    } else {
        for (uint32_t k = 0; k < head; k += 4) *reinterpret_cast<uint32_t *>(dst + k) = *reinterpret_cast<const uint32_t *>(src + k);
        for (uint32_t k = head + body; k < head + body + tail; k += 4)
            *reinterpret_cast<uint32_t *>(dst + k) = *reinterpret_cast<const uint32_t *>(src + k);
    }
    return body;
}

// Stage `nrows` row chunks of `len` elements each (element size es, global row stride row_stride elements).
// flat: the rows are back to back in global memory (row_stride == len) -> a single span.
__device__ __forceinline__ uint32_t stage_rows(char *region, const char *g0, long long row_stride, int nrows, int len, int es,
                                               int pitch, bool flat, uint64_t *bar, int lane) {
    uint32_t tx = 0;
    if (flat) {
        if (lane == 0) tx = stage_span(region, g0, (uint32_t)nrows * len * es, bar, es);
    } else {
        for (int r = lane; r < nrows; r += 32)
            tx += stage_span(region + (size_t)r * pitch, g0 + (size_t)r * row_stride * es, (uint32_t)len * es, bar, es);
    }
    return tx;
}

// Consumer side: shared-memory address of element 0 of row chunk r inside a region staged by stage_rows.
__device__ __forceinline__ const char *staged_row(const char *region, const char *g0, long long row_stride, int r, int len, int es,
                                                  int pitch, bool flat) {
    if (flat) return region + ((uintptr_t)g0 & 15u) + (size_t)r * len * es;
    const char *grow = g0 + (size_t)r * row_stride * es;
    return region + (size_t)r * pitch + ((uintptr_t)grow & 15u);
}

struct ItemCoord { int b, g, tile, row0, nrows; };
__device__ __forceinline__ ItemCoord decode_item(const ScanArgs &a, int item) {
    ItemCoord c;
    c.tile = item % a.tiles_per_group;
    const int bg = item / a.tiles_per_group;
    c.g = bg % a.G;
    c.b = bg / a.G;
    const int in_group = c.tile * a.RT;
    c.row0 = c.g * a.rows_per_group + in_group;
    c.nrows = min(a.RT, a.rows_per_group - in_group);
    return c;
}

// ------------------------------------------------------------------------------------------------
// Warp-shuffle scans of (a, b) pairs over segments of `lpr` lanes (lpr is a power of two, j = lane % lpr).
// Forward: on return (pa, pb) is the inclusive aggregate of lanes [0..j]; (ea, eb) the exclusive one.
__device__ __forceinline__ void seg_scan_fwd(float &pa, float &pb, float &ea, float &eb, int j, int lpr) {
    for (int off = 1; off < lpr; off <<= 1) {
        float qa = __shfl_up_sync(0xffffffffu, pa, off, lpr);
        float qb = __shfl_up_sync(0xffffffffu, pb, off, lpr);
        const bool ok = j >= off;
        qa = ok ? qa : 1.f;
        qb = ok ? qb : 0.f;
        pb = fmaf(pa, qb, pb);   // earlier segment (qa,qb) then ours (pa,pb): (qa pa, pa qb + pb)
        pa = pa * qa;
    }
    ea = __shfl_up_sync(0xffffffffu, pa, 1, lpr);
    eb = __shfl_up_sync(0xffffffffu, pb, 1, lpr);
    if (j == 0) { ea = 1.f; eb = 0.f; }
}
// Reverse (suffix) scan: G_t = rb_t + ra_t * G_{t+1}.  On return (pa, pb) aggregates lanes [j..lpr-1],
// (ea, eb) lanes [j+1..lpr-1].
__device__ __forceinline__ void seg_scan_rev(float &pa, float &pb, float &ea, float &eb, int j, int lpr) {
    for (int off = 1; off < lpr; off <<= 1) {
        float qa = __shfl_down_sync(0xffffffffu, pa, off, lpr);
        float qb = __shfl_down_sync(0xffffffffu, pb, off, lpr);
        const bool ok = j + off < lpr;
        qa = ok ? qa : 1.f;
        qb = ok ? qb : 0.f;
        pb = fmaf(pa, qb, pb);
        pa = pa * qa;
    }
    ea = __shfl_down_sync(0xffffffffu, pa, 1, lpr);
    eb = __shfl_down_sync(0xffffffffu, pb, 1, lpr);
    if (j == lpr - 1) { ea = 1.f; eb = 0.f; }
}

__device__ __forceinline__ float seg_sum(float v, int lpr) {
    for (int off = lpr >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off, lpr);
    return v;
}

}  // namespace mia
