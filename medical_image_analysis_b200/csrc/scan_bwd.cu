// Backward selective scan for sm_100a.
//
// Replaces selective_scan_bwd_kernel of the reference
// (R2GenCSR/VMamba/kernels/selective_scan/csrc/selective_scan/cusoflex/selective_scan_bwd_kernel_oflex.cuh:74-291).
// Gradient algebra (same quantities as lines 216-224 / 245-259 there, re-derived):
//   h_t  = a_t h_{t-1} + b_t,  a_t = exp(dl_t A),  b_t = dl_t u_t B_t,  y_t = sum_n C_t h_t + D u_t
//   G_t  = a_t (dy_t C_t + G_{t+1})              (suffix scan; G_{L} = 0)
//   g_t  = dy_t C_t + G_{t+1}                     (= dL/dh_t)
//   du_t = D dy_t + sum_n g_t B_t dl_t            ddl_t = sum_n g_t (B_t u_t + A a_t h_{t-1})
//   dA   = sum_t g_t dl_t a_t h_{t-1}             dB_t = sum_rows g_t dl_t u_t      dC_t = sum_rows dy_t h_t
//   ddelta_t = ddl_t * sigmoid(delta_t + bias) when softplus;  dD = sum dy u;  dbias = sum ddelta
// Chunks are walked last-to-first; the forward state at a chunk start comes from the checkpoints x written by
// the forward kernel, the suffix value G crosses chunks through a shared-memory carry.
//
// Reductions are deterministic: dB/dC over the rows of a group are summed in registers across the rows a warp
// owns, then across warps through shared memory, and leave the CTA as per-tile partials (d_state <= 2, the
// reference's shipped configuration); dA/dD/dbias leave as per-(batch,row) partials.  A finalize kernel
// (scan_api.cu) folds the partials.  For d_state > 2 dB/dC use fp32 vector reductions into an L2-resident
// accumulator (red.global.add.v4.f32), like the reference's atomics.
#include "scan_common.cuh"

namespace mia {

constexpr int kAccN = 2;  // dB/dC register accumulation for d_state <= kAccN

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void consumer_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

template <typename T>
__global__ void __launch_bounds__(512, 1) ss_bwd_kernel(const __grid_constant__ ScanArgs a) {
    extern __shared__ __align__(128) char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NW = a.n_consumer_warps;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + a.off_bars);
    uint64_t *empty = full + a.stages;
    const int N = a.N, L = a.L, CH = a.CH, RT = a.RT;
    float *carryG = reinterpret_cast<float *>(smem + a.off_carry);  // [RT][N] suffix value entering from the next chunk
    float *carryA = carryG + RT * N;                                // [RT][N] dA accumulated over chunks
    float *carryD = carryA + RT * N;                                // [RT]
    float *carryBias = carryD + RT;                                 // [RT]
    float *red = reinterpret_cast<float *>(smem + a.off_red);       // [NW][256]
    constexpr int es = (int)sizeof(T);
    const int eso = a.out_f32 ? 4 : es;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, NW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NW) {
        // ===================== producer warp =====================
        int k = 0;
        for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
            const ItemCoord ic = decode_item(a, item);
            const int dg0 = ic.row0 / a.delta_ratio;
            const int ndrows = (ic.row0 + ic.nrows - 1) / a.delta_ratio - dg0 + 1;
            for (int c = a.n_chunks - 1; c >= 0; --c, ++k) {
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
                const char *gdo = (const char *)a.dout + ((size_t)ic.b * a.dout_bs + (size_t)ic.row0 * a.dout_ds + l0) * eso;
                tx += stage_rows(st + a.off_dout, gdo, a.dout_ds, ic.nrows, len, eso, a.rowo_pitch, a.flat_dout, full + s, lane);
                if (a.has_z) {
                    const char *gz = (const char *)a.z + ((size_t)ic.b * a.z_bs + (size_t)ic.row0 * a.z_ds + l0) * es;
                    tx += stage_rows(st + a.off_z, gz, a.z_ds, ic.nrows, len, es, a.row_pitch, a.flat_z, full + s, lane);
                    const char *gos = (const char *)a.out_saved + ((size_t)ic.b * a.osaved_bs + (size_t)ic.row0 * a.osaved_ds + l0) * eso;
                    tx += stage_rows(st + a.off_osaved, gos, a.osaved_ds, ic.nrows, len, eso, a.rowo_pitch, a.flat_osaved, full + s, lane);
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
        // ===================== consumer warps =====================
        const int LPR = a.LPR, RPP = 32 / LPR, sub = lane / LPR, j = lane % LPR;
        const int tok0 = j * kTok;
        const float *Ap = reinterpret_cast<const float *>(a.A);
        const float *Dp = reinterpret_cast<const float *>(a.D);
        const float *biasp = reinterpret_cast<const float *>(a.delta_bias);
        const bool acc_regs = !a.bc_atomic;
        const int Lp = (L + 3) & ~3;  // row pitch of the atomic dB/dC accumulators
        int k = 0;
        for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
            const ItemCoord ic = decode_item(a, item);
            const int dg0 = ic.row0 / a.delta_ratio;
            for (int c = a.n_chunks - 1; c >= 0; --c, ++k) {
                const int s = k % a.stages;
                mbar_wait(full + s, (k / a.stages) & 1);
                const char *st = smem + (size_t)s * a.stage_bytes;
                const int l0 = c * CH, len = min(CH, L - l0);
                const int nval = max(0, min(kTok, len - tok0));
                const bool last_chunk = c == a.n_chunks - 1, first_chunk = c == 0;
                const char *gu = (const char *)a.u + ((size_t)ic.b * a.u_bs + (size_t)ic.row0 * a.u_ds + l0) * es;
                const char *gd = (const char *)a.delta + ((size_t)ic.b * a.delta_bs + (size_t)dg0 * a.delta_ds + l0) * es;
                const char *gdo = (const char *)a.dout + ((size_t)ic.b * a.dout_bs + (size_t)ic.row0 * a.dout_ds + l0) * eso;
                const char *gz = a.has_z ? (const char *)a.z + ((size_t)ic.b * a.z_bs + (size_t)ic.row0 * a.z_ds + l0) * es : nullptr;
                const char *gos = a.has_z ? (const char *)a.out_saved + ((size_t)ic.b * a.osaved_bs + (size_t)ic.row0 * a.osaved_ds + l0) * eso : nullptr;
                const char *gB = (const char *)a.B + ((size_t)ic.b * a.B_bs + (size_t)ic.g * a.B_gs + l0) * es;
                const char *gC = (const char *)a.C + ((size_t)ic.b * a.C_bs + (size_t)ic.g * a.C_gs + l0) * es;

                float dBacc[kAccN][kTok], dCacc[kAccN][kTok];
#pragma unroll
                for (int n = 0; n < kAccN; ++n)
#pragma unroll
                    for (int i = 0; i < kTok; ++i) dBacc[n][i] = dCacc[n][i] = 0.f;

                for (int rbase = warp * RPP; rbase < ic.nrows; rbase += NW * RPP) {
                    const bool active = rbase + sub < ic.nrows;
                    const int r = active ? rbase + sub : ic.nrows - 1;
                    const int d = ic.row0 + r;
                    const int dgrp = d / a.delta_ratio;
                    float u8[kTok], dl[kTok], dy[kTok], du[kTok], ddl[kTok], sg[kTok];
                    ld8<T>(staged_row(st + a.off_u, gu, a.u_ds, r, len, es, a.row_pitch, a.flat_u) + tok0 * es, u8);
                    ld8<T>(staged_row(st + a.off_delta, gd, a.delta_ds, dgrp - dg0, len, es, a.row_pitch, a.flat_delta) + tok0 * es, dl);
                    {
                        const char *p = staged_row(st + a.off_dout, gdo, a.dout_ds, r, len, eso, a.rowo_pitch, a.flat_dout) + tok0 * eso;
                        if (a.out_f32) ld8<float>(p, dy); else ld8<T>(p, dy);
                    }
                    const float bias = biasp ? __ldg(biasp + dgrp) : 0.f;
                    const float Dv = Dp ? __ldg(Dp + d) : 0.f;
                    if (a.has_z) {
                        // out_z = y * silu(z):  dy = dout * silu(z);  dz = dout * y * sigmoid(z) * (1 + z (1 - sigmoid(z)))
                        float z8[kTok], o8[kTok], dz[kTok];
                        ld8<T>(staged_row(st + a.off_z, gz, a.z_ds, r, len, es, a.row_pitch, a.flat_z) + tok0 * es, z8);
                        const char *p = staged_row(st + a.off_osaved, gos, a.osaved_ds, r, len, eso, a.rowo_pitch, a.flat_osaved) + tok0 * eso;
                        if (a.out_f32) ld8<float>(p, o8); else ld8<T>(p, o8);
#pragma unroll
                        for (int i = 0; i < kTok; ++i) {
                            const float sz = rcpf(1.f + ex2f(-z8[i] * kLog2e));
                            dz[i] = dy[i] * o8[i] * sz * (1.f + z8[i] * (1.f - sz));
                            dy[i] = dy[i] * z8[i] * sz;
                        }
                        if (active && nval > 0) {
                            char *gdz = (char *)a.dz + ((size_t)ic.b * a.dz_bs + (size_t)d * a.dz_ds + l0 + tok0) * es;
                            st8<T>(gdz, dz, nval);
                        }
                    }
                    float dDv = 0.f;
#pragma unroll
                    for (int i = 0; i < kTok; ++i) {
                        const bool valid = i < nval;
                        float v = dl[i] + bias;
                        sg[i] = 1.f;
                        if (a.softplus) v = softplus_f<true>(v, sg[i]);
                        dl[i] = valid ? v : 0.f;
                        u8[i] = valid ? u8[i] : 0.f;
                        dy[i] = valid ? dy[i] : 0.f;
                        du[i] = Dv * dy[i];
                        dDv = fmaf(dy[i], u8[i], dDv);
                        ddl[i] = 0.f;
                    }
                    for (int n = 0; n < N; ++n) {
                        const float Araw = __ldg(Ap + (size_t)d * a.A_ds + (size_t)n * a.A_ns);
                        const float A2 = Araw * kLog2e;
                        float B8[kTok], C8[kTok];
                        ld8<T>(staged_row(st + a.off_B, gB, a.B_ns, n, len, es, a.bc_pitch, a.flat_B) + tok0 * es, B8);
                        ld8<T>(staged_row(st + a.off_C, gC, a.C_ns, n, len, es, a.bc_pitch, a.flat_C) + tok0 * es, C8);
                        if (nval < kTok) {
#pragma unroll
                            for (int i = 0; i < kTok; ++i) {
                                B8[i] = i < nval ? B8[i] : 0.f;
                                C8[i] = i < nval ? C8[i] : 0.f;
                            }
                        }
                        // ---- forward recompute: lane aggregate, warp scan, per-token states
                        float av[kTok], ah[kTok];
                        float pa = 1.f, pb = 0.f;
#pragma unroll
                        for (int i = 0; i < kTok; ++i) {
                            av[i] = ex2f(dl[i] * A2);
                            ah[i] = dl[i] * u8[i] * B8[i];      // b_t for now
                            pb = fmaf(av[i], pb, ah[i]);
                            pa *= av[i];
                        }
                        float h0 = 0.f;
                        if (!first_chunk) h0 = __ldg(a.x + (((size_t)(ic.b * a.dim + d) * a.n_chunks + (c - 1)) * N + n) * 2 + 1);
                        float ea, eb;
                        seg_scan_fwd(pa, pb, ea, eb, j, LPR);
                        float hm = fmaf(ea, h0, eb);
                        float dCv[kTok];
                        float ra = 1.f, rb = 0.f;
#pragma unroll
                        for (int i = 0; i < kTok; ++i) {
                            const float t = av[i] * hm;          // a_t h_{t-1}
                            hm = t + ah[i];                      // h_t
                            ah[i] = t;
                            dCv[i] = dy[i] * hm;
                        }
                        // ---- suffix scan of G
#pragma unroll
                        for (int i = kTok - 1; i >= 0; --i) {
                            rb = av[i] * fmaf(dy[i], C8[i], rb);
                            ra *= av[i];
                        }
                        float gin = 0.f;
                        if (!last_chunk) gin = carryG[r * N + n];
                        seg_scan_rev(ra, rb, ea, eb, j, LPR);
                        float Gn = fmaf(ea, gin, eb);            // G entering from the first token after this lane
                        float dAv = 0.f;
                        float dBv[kTok];
#pragma unroll
                        for (int i = kTok - 1; i >= 0; --i) {
                            const float g = fmaf(dy[i], C8[i], Gn);
                            const float gB = g * B8[i];
                            du[i] = fmaf(gB, dl[i], du[i]);
                            const float gah = g * ah[i];
                            ddl[i] = fmaf(gB, u8[i], ddl[i]);
                            ddl[i] = fmaf(gah, Araw, ddl[i]);
                            dAv = fmaf(gah, dl[i], dAv);
                            dBv[i] = g * dl[i] * u8[i];
                            Gn = av[i] * g;
                        }
                        __syncwarp();
                        if (j == 0 && active) carryG[r * N + n] = Gn;   // G at the first token of this chunk, for chunk c-1
                        // ---- dA: reduce over the row's lanes, accumulate over chunks
                        dAv = seg_sum(dAv, LPR);
                        if (j == 0 && active) {
                            const float tot = last_chunk ? dAv : dAv + carryA[r * N + n];
                            if (first_chunk) a.part_dA[(size_t)(ic.b * a.dim + d) * N + n] = tot;
                            else carryA[r * N + n] = tot;
                        }
                        // ---- dB / dC: reduce over rows
                        if (acc_regs) {
                            if (active) {
#pragma unroll
                                for (int m = 0; m < kAccN; ++m)
                                    if (m == n) {
#pragma unroll
                                        for (int i = 0; i < kTok; ++i) { dBacc[m][i] += dBv[i]; dCacc[m][i] += dCv[i]; }
                                    }
                            }
                        } else if (active && nval > 0) {
                            float *pB = a.acc_dB + ((size_t)(ic.b * a.G + ic.g) * N + n) * Lp + l0 + tok0;
                            float *pC = a.acc_dC + ((size_t)(ic.b * a.G + ic.g) * N + n) * Lp + l0 + tok0;
                            red_add_v4(pB, dBv[0], dBv[1], dBv[2], dBv[3]);
                            red_add_v4(pC, dCv[0], dCv[1], dCv[2], dCv[3]);
                            if (nval > 4) {
                                red_add_v4(pB + 4, dBv[4], dBv[5], dBv[6], dBv[7]);
                                red_add_v4(pC + 4, dCv[4], dCv[5], dCv[6], dCv[7]);
                            }
                        }
                    }
                    // ---- per-row epilogue
                    float dbv = 0.f;
#pragma unroll
                    for (int i = 0; i < kTok; ++i) {
                        ddl[i] *= sg[i];
                        dbv += ddl[i];
                    }
                    if (active && nval > 0) {
                        char *gdu = (char *)a.du + ((size_t)ic.b * a.du_bs + (size_t)d * a.du_ds + l0 + tok0) * es;
                        st8<T>(gdu, du, nval);
                        if (a.delta_ratio == 1) {
                            char *gdd = (char *)a.ddelta + ((size_t)ic.b * a.dd_bs + (size_t)d * a.dd_ds + l0 + tok0) * es;
                            st8<T>(gdd, ddl, nval);
                        } else {
                            float *gdd = a.ddelta_full + ((size_t)(ic.b * a.dim + d)) * L + l0 + tok0;
                            st8<float>(gdd, ddl, nval);
                        }
                    }
                    dDv = seg_sum(dDv, LPR);
                    dbv = seg_sum(dbv, LPR);
                    if (j == 0 && active) {
                        const float tD = last_chunk ? dDv : dDv + carryD[r];
                        const float tb = last_chunk ? dbv : dbv + carryBias[r];
                        if (first_chunk) {
                            a.part_dD[(size_t)ic.b * a.dim + d] = tD;
                            a.part_dbias[(size_t)ic.b * a.dim + d] = tb;
                        } else {
                            carryD[r] = tD;
                            carryBias[r] = tb;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s);

                if (acc_regs) {
                    // ---- fold this warp's dB/dC (over its rows) with the other warps', emit the tile partial.
                    // Slot p = lane*8 + i = sub*CH + token.  Fixed summation order -> bit-reproducible.
                    const int tid = threadIdx.x;  // consumer threads are 0 .. NW*32-1
                    for (int n = 0; n < N; ++n) {
#pragma unroll
                        for (int which = 0; which < 2; ++which) {
                            float *mine = red + warp * 256 + lane * kTok;
#pragma unroll
                            for (int m = 0; m < kAccN; ++m)
                                if (m == n) {
#pragma unroll
                                    for (int i = 0; i < kTok; ++i) mine[i] = which == 0 ? dBacc[m][i] : dCacc[m][i];
                                }
                            consumer_bar(NW * 32);
                            if (tid < len) {
                                float sum = 0.f;
                                for (int w = 0; w < NW; ++w)
                                    for (int sb = 0; sb < RPP; ++sb) sum += red[w * 256 + sb * CH + tid];
                                float *dst = (which == 0 ? a.acc_dB : a.acc_dC) + ((size_t)item * N + n) * L + l0 + tid;
                                *dst = sum;
                            }
                            consumer_bar(NW * 32);
                        }
                    }
                }
            }
        }
    }
}

template <typename T>
cudaError_t launch_bwd(const ScanArgs &a, int grid, cudaStream_t stream) {
    auto kernel = &ss_bwd_kernel<T>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
    if (e != cudaSuccess) return e;
    kernel<<<grid, (a.n_consumer_warps + 1) * 32, a.smem_bytes, stream>>>(a);
    return cudaGetLastError();
}

template cudaError_t launch_bwd<float>(const ScanArgs &, int, cudaStream_t);
template cudaError_t launch_bwd<__half>(const ScanArgs &, int, cudaStream_t);
template cudaError_t launch_bwd<__nv_bfloat16>(const ScanArgs &, int, cudaStream_t);

}  // namespace mia
