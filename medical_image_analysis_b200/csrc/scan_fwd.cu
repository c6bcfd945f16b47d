// Forward selective scan for sm_100a.
//
// Replaces selective_scan_fwd_kernel of the reference
// (R2GenCSR/VMamba/kernels/selective_scan/csrc/selective_scan/cusoflex/selective_scan_fwd_kernel_oflex.cuh:68-181):
//   delta' = softplus(delta + bias);  h_n[l] = exp(delta' A_n) h_n[l-1] + delta' u B_n[l];
//   y[l]   = sum_n C_n[l] h_n[l] + D u[l];   (optional) out_z = y * silu(z)
// and writes the per-chunk (prod a, h) checkpoints x that the backward restarts from.
//
// Not a port: instead of one CTA per (batch, row) with CUB block loads/scans and 2 __syncthreads per state,
// a persistent CTA walks (batch, group, row-tile) items; a producer warp TMA-stages whole tiles (B/C once per
// tile, not once per row) through an mbarrier ring, and each consumer warp scans its rows with register-serial
// 8-token segments + one warp-shuffle scan per state.  No block-wide synchronisation in the steady state.
#include "scan_common.cuh"

namespace mia {

template <typename T>
__global__ void __launch_bounds__(512, 1) ss_fwd_kernel(const __grid_constant__ ScanArgs a) {
    extern __shared__ __align__(128) char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NW = a.n_consumer_warps;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + a.off_bars);
    uint64_t *empty = full + a.stages;
    float2 *carry = reinterpret_cast<float2 *>(smem + a.off_carry);  // [RT][N] running (prod a, h)
    constexpr int es = (int)sizeof(T);

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, NW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int N = a.N, L = a.L, CH = a.CH;

    if (warp == NW) {
        // ===================== producer warp: TMA-stage tiles =====================
        int k = 0;
        for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
            const ItemCoord ic = decode_item(a, item);
            const int dg0 = ic.row0 / a.delta_ratio;
            const int ndrows = (ic.row0 + ic.nrows - 1) / a.delta_ratio - dg0 + 1;
            for (int c = 0; c < a.n_chunks; ++c, ++k) {
                const int s = k % a.stages;
                const int use = k / a.stages;
                if (use > 0) mbar_wait(empty + s, (use - 1) & 1);
                char *st = smem + (size_t)s * a.stage_bytes;
                const int l0 = c * CH, len = min(CH, L - l0);
                uint32_t tx = 0;
                const char *gu = (const char *)a.u + ((size_t)ic.b * a.u_bs + (size_t)ic.row0 * a.u_ds + l0) * es;
                tx += stage_rows(st + a.off_u, gu, a.u_ds, ic.nrows, len, es, a.row_pitch, a.flat_u, full + s, lane);
                const char *gd = (const char *)a.delta + ((size_t)ic.b * a.delta_bs + (size_t)dg0 * a.delta_ds + l0) * es;
                tx += stage_rows(st + a.off_delta, gd, a.delta_ds, ndrows, len, es, a.row_pitch, a.flat_delta, full + s, lane);
                if (a.has_z) {
                    const char *gz = (const char *)a.z + ((size_t)ic.b * a.z_bs + (size_t)ic.row0 * a.z_ds + l0) * es;
                    tx += stage_rows(st + a.off_z, gz, a.z_ds, ic.nrows, len, es, a.row_pitch, a.flat_z, full + s, lane);
                }
                const char *gB = (const char *)a.B + ((size_t)ic.b * a.B_bs + (size_t)ic.g * a.B_gs + l0) * es;
                tx += stage_rows(st + a.off_B, gB, a.B_ns, N, len, es, a.bc_pitch, a.flat_B, full + s, lane);
                const char *gC = (const char *)a.C + ((size_t)ic.b * a.C_bs + (size_t)ic.g * a.C_gs + l0) * es;
                tx += stage_rows(st + a.off_C, gC, a.C_ns, N, len, es, a.bc_pitch, a.flat_C, full + s, lane);
                tx = __reduce_add_sync(0xffffffffu, tx);
                if (lane == 0) mbar_arrive_expect_tx(full + s, tx);
            }
        }
    } else if (warp < NW) {
        // ===================== consumer warps: scan rows =====================
        const int LPR = a.LPR, RPP = 32 / LPR, sub = lane / LPR, j = lane % LPR;
        const int tok0 = j * kTok;
        const float *Ap = reinterpret_cast<const float *>(a.A);
        const float *Dp = reinterpret_cast<const float *>(a.D);
        const float *biasp = reinterpret_cast<const float *>(a.delta_bias);
        int k = 0;
        for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
            const ItemCoord ic = decode_item(a, item);
            const int dg0 = ic.row0 / a.delta_ratio;
            for (int c = 0; c < a.n_chunks; ++c, ++k) {
                const int s = k % a.stages;
                mbar_wait(full + s, (k / a.stages) & 1);
                const char *st = smem + (size_t)s * a.stage_bytes;
                const int l0 = c * CH, len = min(CH, L - l0);
                const int nval = max(0, min(kTok, len - tok0));
                const char *gu = (const char *)a.u + ((size_t)ic.b * a.u_bs + (size_t)ic.row0 * a.u_ds + l0) * es;
                const char *gd = (const char *)a.delta + ((size_t)ic.b * a.delta_bs + (size_t)dg0 * a.delta_ds + l0) * es;
                const char *gz = a.has_z ? (const char *)a.z + ((size_t)ic.b * a.z_bs + (size_t)ic.row0 * a.z_ds + l0) * es : nullptr;
                const char *gB = (const char *)a.B + ((size_t)ic.b * a.B_bs + (size_t)ic.g * a.B_gs + l0) * es;
                const char *gC = (const char *)a.C + ((size_t)ic.b * a.C_bs + (size_t)ic.g * a.C_gs + l0) * es;

                for (int rbase = warp * RPP; rbase < ic.nrows; rbase += NW * RPP) {
                    const bool active = rbase + sub < ic.nrows;
                    const int r = active ? rbase + sub : ic.nrows - 1;  // idle sub-rows shadow a valid row, never store
                    const int d = ic.row0 + r;
                    const int dgrp = d / a.delta_ratio;
                    float u8[kTok], dl[kTok], y[kTok];
                    ld8<T>(staged_row(st + a.off_u, gu, a.u_ds, r, len, es, a.row_pitch, a.flat_u) + tok0 * es, u8);
                    ld8<T>(staged_row(st + a.off_delta, gd, a.delta_ds, dgrp - dg0, len, es, a.row_pitch, a.flat_delta) + tok0 * es, dl);
                    const float bias = biasp ? __ldg(biasp + dgrp) : 0.f;
                    const float Dv = Dp ? __ldg(Dp + d) : 0.f;
#pragma unroll
                    for (int i = 0; i < kTok; ++i) {
                        const bool valid = i < nval;
                        float sg;
                        float v = dl[i] + bias;
                        if (a.softplus) v = softplus_f<false>(v, sg);
                        dl[i] = valid ? v : 0.f;            // a = exp2(0) = 1 for padding tokens
                        const float uu = valid ? u8[i] : 0.f;
                        y[i] = Dv * uu;
                        u8[i] = dl[i] * uu;                 // from here on u8 holds delta' * u
                    }
                    for (int n = 0; n < N; ++n) {
                        const float A2 = __ldg(Ap + (size_t)d * a.A_ds + (size_t)n * a.A_ns) * kLog2e;
                        float B8[kTok], C8[kTok];
                        ld8<T>(staged_row(st + a.off_B, gB, a.B_ns, n, len, es, a.bc_pitch, a.flat_B) + tok0 * es, B8);
                        ld8<T>(staged_row(st + a.off_C, gC, a.C_ns, n, len, es, a.bc_pitch, a.flat_C) + tok0 * es, C8);
                        if (nval < kTok) {
#pragma unroll
                            for (int i = 0; i < kTok; ++i) B8[i] = i < nval ? B8[i] : 0.f;  // staged tail bytes are undefined
                        }
                        float2 cin = make_float2(1.f, 0.f);
                        if (c > 0) cin = carry[r * N + n];
                        float av[kTok], bv[kTok];
                        float pa = 1.f, pb = 0.f;
#pragma unroll
                        for (int i = 0; i < kTok; ++i) {
                            av[i] = ex2f(dl[i] * A2);
                            bv[i] = u8[i] * B8[i];
                            pb = fmaf(av[i], pb, bv[i]);
                            pa *= av[i];
                        }
                        float ea, eb;
                        seg_scan_fwd(pa, pb, ea, eb, j, LPR);
                        float h = fmaf(ea, cin.y, eb);
#pragma unroll
                        for (int i = 0; i < kTok; ++i) {
                            h = fmaf(av[i], h, bv[i]);
                            y[i] = fmaf(h, C8[i], y[i]);
                        }
                        __syncwarp();
                        if (j == LPR - 1 && active) {
                            const float2 cout = make_float2(pa * cin.x, fmaf(pa, cin.y, pb));
                            carry[r * N + n] = cout;
                            // checkpoint layout of the reference: x[b][d][chunk][2n] = prod a, [2n+1] = h  (fwd kernel :164-167)
                            float2 *xp = reinterpret_cast<float2 *>(a.x) + ((size_t)(ic.b * a.dim + d) * a.n_chunks + c) * N + n;
                            *xp = cout;
                        }
                    }
                    if (active && nval > 0) {
                        const size_t eo = a.out_f32 ? 4 : es;
                        char *go = (char *)a.out + ((size_t)ic.b * a.out_bs + (size_t)d * a.out_ds + l0 + tok0) * eo;
                        if (a.out_f32) st8<float>(go, y, nval); else st8<T>(go, y, nval);
                        if (a.has_z) {
                            float z8[kTok];
                            ld8<T>(staged_row(st + a.off_z, gz, a.z_ds, r, len, es, a.row_pitch, a.flat_z) + tok0 * es, z8);
#pragma unroll
                            for (int i = 0; i < kTok; ++i) {
                                const float sg = rcpf(1.f + ex2f(-z8[i] * kLog2e));
                                y[i] = y[i] * z8[i] * sg;
                            }
                            char *gz_out = (char *)a.out_z + ((size_t)ic.b * a.outz_bs + (size_t)d * a.outz_ds + l0 + tok0) * eo;
                            if (a.out_f32) st8<float>(gz_out, y, nval); else st8<T>(gz_out, y, nval);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s);
            }
        }
    }
}

template <typename T>
cudaError_t launch_fwd(const ScanArgs &a, int grid, cudaStream_t stream) {
    auto kernel = &ss_fwd_kernel<T>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
    if (e != cudaSuccess) return e;
    kernel<<<grid, (a.n_consumer_warps + 1) * 32, a.smem_bytes, stream>>>(a);
    return cudaGetLastError();
}

template cudaError_t launch_fwd<float>(const ScanArgs &, int, cudaStream_t);
template cudaError_t launch_fwd<__half>(const ScanArgs &, int, cudaStream_t);
template cudaError_t launch_fwd<__nv_bfloat16>(const ScanArgs &, int, cudaStream_t);

}  // namespace mia
