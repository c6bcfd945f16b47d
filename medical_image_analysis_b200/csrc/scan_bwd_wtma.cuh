// Backward selective scan, WINDOWED row-serial path (d_state == 1) fed by 2-D TENSOR-MAP TMA (cp.async.bulk.tensor.2d, SASS
// UTMALDG / UTMASTG) on forward-provided block states.  Same algorithm as scan_bwd_win.cuh (lane per row, phase 2 only, windows
// of 32 tokens from the end of the row, G carried in a register); what changes is the data movement:
//   * a window tile [32 rows x 32 tokens] of u / delta / dout is ONE box copy per tensor issued by one lane (the cp.async version
//     spends ~170 instructions per window on 8-byte pieces and their addresses), completion on an mbarrier;
//   * du / ddelta are written into a separate output tile and leave with one TMA box store each, so the input stage can be
//     refilled without waiting for the stores to drain (the output tile itself is single: its stores get a recompute phase,
//     ~a third of a block, to be read out before the next window's first write);
//   * rows whose byte pitch is not a multiple of 16 (L = 196 bf16 elements = 392 bytes) cannot be a tensor-map dimension --
//     two consecutive rows can (784 bytes): the tensor is mapped as [rows / 2][2 L] and a tile is two boxes, one at inner
//     coordinate t0 (even rows) and one at L + t0 (odd rows); lane l then owns row 2 (l % 16) + l / 16 of the 32-row batch.
//     A window that crosses the end of the row would, for the even rows, run into the next row: it is read (harmlessly, the
//     ragged block never touches those tokens) but its outputs are stored by plain coalesced stores, not by a box store;
//   * tiles carry the 128-byte TMA swizzle: with dense 64- / 128-byte tile rows the 32 lanes (one row each, same token) would
//     otherwise collide on 2-4 banks; swizzled they spread over all eight 16-byte slots (4 lanes per slot).
// Preconditions (host-checked): as scan_bwd_win.cuh, plus L * es % 8 == 0 (always true for L % 4 == 0), 16-byte aligned tensors.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "scan_bwd_rows.cuh"
#include "scan_fwd_cw.cuh"

namespace mia {

constexpr int kWtTok = 32;       // tokens per window

struct WinTmaArgs {
    int batch, dim, L, G, rows_per_group;
    int softplus;
    int n_items, nblk, nwin, g;             // g: rows per tensor-map row (1 or 2)
    int in_stage, out_stage;                // bytes of one input stage (u, delta, dout tiles) / output stage (du, ddelta tiles)
    int off_d, off_o, off_outs, off_bc32, off_bar, smem_bytes;
    const void *A, *B, *C, *D, *delta_bias;
    const float *hblk;
    void *du, *ddelta;                      // for the plain stores of a ragged last window (g == 2)
    float *part_dA, *part_dD, *part_dbias, *acc_dB, *acc_dC;
    long long B_bs, B_gs, C_bs, C_gs;
};

// One 16-token block (swizzled tiles): recompute from the block's entering state, suffix recurrence, du / ddelta into the output
// tile, dB / dC butterfly.  tb: first token of the block, wb: byte offset of the block inside the window row (in dtype).
template <typename T, typename TO, bool kSoftplus, bool kFull>
__device__ __forceinline__ void bwd_block_sw(const bool wait_out, const int tb, const int nq_in, const int tend, const int lane, const float h0, const char *tu,
                                             const char *td, const char *to, char *tdu, char *tdd, const SwzRow ri, const SwzRow ro, const int wbi,
                                             const int wbo, const float *Bf, const float *Cf, float *accB, float *accC, const float2 bl2,
                                             const float2 A2, const float2 Aln2, const float2 D2, float &G, float2 &dA2, float2 &dD2, float2 &db2) {
    constexpr int es = (int)sizeof(T), eo = (int)sizeof(TO);
    const int nq = kFull ? 4 : nq_in;
    BlkRegs R;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    float h = h0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (kFull || q < nq)
            recompute_quad<T, kSoftplus>(R, q, tu + ri.at(wbi + 4 * q * es), td + ri.at(wbi + 4 * q * es), Bf + tb + 4 * q, h, bl2, A2, Aln2);
    if (wait_out) {
        // first block of a window: the (single) output tile is about to be overwritten -- the previous window's box stores,
        // issued a recompute phase ago, must have been read out of shared memory
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
    }
#pragma unroll
    for (int q = 3; q >= 0; --q) {
        if (kFull || q < nq) {
            const int t = tb + 4 * q;
            float2 dy[2], Cv[2], du[2], dd[2], uu[2];
            Quad<TO>::ld(to + ro.at(wbo + 4 * q * eo), dy);
            Quad<float>::ld(reinterpret_cast<const char *>(Cf + t), Cv);
            Quad<T>::ld(tu + ri.at(wbi + 4 * q * es), uu);      // the input tile is not overwritten here: u is re-read, not kept in registers
#pragma unroll
            for (int p = 1; p >= 0; --p) {
                const int k = 2 * q + p;
                const float2 pc = mul2(dy[p], Cv[p]);
                const float2 ap = mul2(R.a[k], pc);
                float2 gg;
                gg.y = pc.y + G; G = fmaf(R.a[k].y, G, ap.y);
                gg.x = pc.x + G; G = fmaf(R.a[k].x, G, ap.x);
                const float2 dBv = mul2(gg, R.e[k]), dCv = mul2(dy[p], R.h[k]);
                v[2 * k] = dBv.x; v[2 * k + 1] = dBv.y;
                v[16 + 2 * k] = dCv.x; v[17 + 2 * k] = dCv.y;
                du[p] = fma2(gg, R.f[k], mul2(dy[p], D2));
                dd[p] = mul2(gg, R.r[k]);
                db2 = add2(db2, dd[p]);
                dA2 = fma2(gg, R.w[k], dA2);
                dD2 = fma2(dy[p], uu[p], dD2);
            }
            Quad<T>::st(tdu + ri.at(wbi + 4 * q * es), du);
            Quad<T>::st(tdd + ri.at(wbi + 4 * q * es), dd);
        }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    const int tt = tb + (lane & 15);
    if (kFull || tt < tend) {
        if (lane < 16) accB[tt] = v[0] * kLn2; else accC[tt] = v[0];
    }
}

template <typename T, bool kSoftplus, bool kOutF32, int kG>
__global__ void __launch_bounds__(32, 12) ss_bwd_wtma_kernel(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_d,
                                                             const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_du,
                                                             const __grid_constant__ CUtensorMap tm_dd, const __grid_constant__ WinTmaArgs a) {
    extern __shared__ char smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array (a cast through an integer would make every tile
    // access a generic LD / ST instead of LDS / STS)
    char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int es = (int)sizeof(T);
    constexpr int eo = kOutF32 ? 4 : es;
    using TO = typename std::conditional<kOutF32, float, T>::type;
    using raw = typename Cvt<T>::raw;
    constexpr int RBi = kWtTok * es, RBo = kWtTok * eo;                  // tile row bytes
    constexpr int kTileI = 32 * RBi, kTileO = 32 * RBo;
    const int lane = threadIdx.x;
    float *Bw = reinterpret_cast<float *>(smem + a.off_bc32), *Cw = Bw + kWtTok;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + a.off_bar);
    if (lane == 0) { mbar_init(full, 1); mbar_init(full + 1, 1); fence_mbar_init(); }
    __syncwarp();
    const int L = a.L, nblk = a.nblk, nwin = a.nwin;
    constexpr int g = kG;
    const int batches_per_group = a.rows_per_group / 32;
    const float *Ap = reinterpret_cast<const float *>(a.A);
    const float *Dp = reinterpret_cast<const float *>(a.D);
    const float *biasp = reinterpret_cast<const float *>(a.delta_bias);
    const SwzRow ri = swz_row<RBi>(lane), ro = swz_row<RBo>(lane);
    // tile row `lane` holds row rho of the 32-row batch: g == 2 -> rows [0, 16) of the tile are the even rows, [16, 32) the odd ones
    const int rho = g == 2 ? 2 * (lane & 15) + (lane >> 4) : lane;
    uint32_t phbits = 0;                                                 // bit s: parity to wait for on full[s]

    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int bt = item % batches_per_group;
        const int bg = item / batches_per_group;
        const int gq = bg % a.G, b = bg / a.G;
        const int row0 = gq * a.rows_per_group + bt * 32;
        const int d = row0 + rho;
        const int srow0 = (b * a.dim + row0) / g;                        // first tensor-map row of the batch
        const raw *gB = reinterpret_cast<const raw *>(a.B) + (size_t)b * a.B_bs + (size_t)gq * a.B_gs;
        const raw *gC = reinterpret_cast<const raw *>(a.C) + (size_t)b * a.C_bs + (size_t)gq * a.C_gs;
        const float *gh = a.hblk + (size_t)item * nblk * 32 + rho;
        float hnext0, hnext1;
        raw bnext = 0, cnext = 0;

        auto load_window = [&](const int w) {
            const int t0 = w * kWtTok;
            if (lane == 0) {
                char *st = smem + (w & 1) * a.in_stage;
                mbar_arrive_expect_tx(full + (w & 1), (uint32_t)(2 * kTileI + kTileO));
#pragma unroll
                for (int s = 0; s < g; ++s) {
                    const int c0 = s * L + t0;
                    tma_box_g2s(st + s * (kTileI / 2), &tm_u, c0, srow0, full + (w & 1));
                    tma_box_g2s(st + a.off_d + s * (kTileI / 2), &tm_d, c0, srow0, full + (w & 1));
                    tma_box_g2s(st + a.off_o + s * (kTileO / 2), &tm_o, c0, srow0, full + (w & 1));
                }
            }
            const int tk = t0 + lane;                                    // this window's B / C element and block states, prefetched into registers
            bnext = tk < L ? __ldg(gB + tk) : (raw)0;
            cnext = tk < L ? __ldg(gC + tk) : (raw)0;
            const int j0 = t0 / kBlk;
            hnext0 = j0 > 0 ? __ldg(gh + j0 * 32) : 0.f;
            hnext1 = j0 + 1 < nblk ? __ldg(gh + (j0 + 1) * 32) : 0.f;
        };
        load_window(nwin - 1);

        const float Araw = __ldg(Ap + d);
        const float Dv = Dp ? __ldg(Dp + d) : 0.f;
        const float2 bl2 = splat2((biasp ? __ldg(biasp + d) : 0.f) * kLog2e), A2 = splat2(Araw), Aln2 = splat2(Araw * kLn2), D2 = splat2(Dv);
        float2 dA2 = make_float2(0.f, 0.f), dD2 = dA2, db2 = dA2;
        float G = 0.f;
        float *accB = a.acc_dB + (size_t)item * L, *accC = a.acc_dC + (size_t)item * L;

        for (int w = nwin - 1; w >= 0; --w) {
            const float hcur0 = hnext0, hcur1 = hnext1;
            const raw bcur = bnext, ccur = cnext;
            if (w > 0) load_window(w - 1);                               // the other input stage was consumed by window w + 1
            const int t0 = w * kWtTok;
            Bw[lane] = Cvt<T>::to_f(bcur) * kLn2;                        // B' = B ln2, C as fp32 (zero past L)
            Cw[lane] = Cvt<T>::to_f(ccur);
            __syncwarp();
            mbar_wait(full + (w & 1), (phbits >> (w & 1)) & 1u);
            phbits ^= 1u << (w & 1);
            const char *st = smem + (w & 1) * a.in_stage;
            char *os = smem + a.off_outs;
            bool first = true;
            const float *Bf = Bw - t0, *Cf = Cw - t0;
#pragma unroll 1
            for (int jb = min(nblk - 1, (t0 + kWtTok) / kBlk - 1); jb >= t0 / kBlk; --jb) {
                const int tb = jb * kBlk, wb = tb - t0;
                const float h0 = jb == t0 / kBlk ? hcur0 : hcur1;
                if (tb + kBlk <= L)
                    bwd_block_sw<T, TO, kSoftplus, true>(first, tb, 4, L, lane, h0, st, st + a.off_d, st + a.off_o, os, os + kTileI, ri, ro, wb * es, wb * eo,
                                                         Bf, Cf, accB, accC, bl2, A2, Aln2, D2, G, dA2, dD2, db2);
                else
                    bwd_block_sw<T, TO, kSoftplus, false>(first, tb, (L - tb) / 4, L, lane, h0, st, st + a.off_d, st + a.off_o, os, os + kTileI, ri, ro,
                                                          wb * es, wb * eo, Bf, Cf, accB, accC, bl2, A2, Aln2, D2, G, dA2, dD2, db2);
                first = false;
            }
            if (g == 2 && t0 + kWtTok > L) {
                // ragged last window of 8-byte-pitch rows: a box store of the even rows would spill into the next row -> plain stores
                __syncwarp();
                for (int i = lane; i < 32 * (kWtTok / 4); i += 32) {     // (tile row, 4-token quad)
                    const int r = i / (kWtTok / 4), q = i % (kWtTok / 4);
                    const int tk = t0 + 4 * q;
                    if (tk < L) {
                        const SwzRow rr = swz_row<RBi>(r);
                        const int arow = 2 * (r & 15) + (r >> 4);
                        const size_t o = (((size_t)b * a.dim + row0 + arow) * L + tk) * es;
                        if (es == 2) {
                            *reinterpret_cast<uint2 *>((char *)a.du + o) = *reinterpret_cast<const uint2 *>(os + rr.at(4 * q * es));
                            *reinterpret_cast<uint2 *>((char *)a.ddelta + o) = *reinterpret_cast<const uint2 *>(os + kTileI + rr.at(4 * q * es));
                        } else {
                            *reinterpret_cast<uint4 *>((char *)a.du + o) = *reinterpret_cast<const uint4 *>(os + rr.at(4 * q * es));
                            *reinterpret_cast<uint4 *>((char *)a.ddelta + o) = *reinterpret_cast<const uint4 *>(os + kTileI + rr.at(4 * q * es));
                        }
                    }
                }
                __syncwarp();
            } else {
                fence_proxy_async();                                     // generic-proxy tile writes -> TMA store
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int s = 0; s < g; ++s) {
                        const int c0 = s * L + t0;
                        tma_box_s2g(&tm_du, os + s * (kTileI / 2), c0, srow0);
                        tma_box_s2g(&tm_dd, os + kTileI + s * (kTileI / 2), c0, srow0);
                    }
                    bulk_commit();
                }
            }
        }
        a.part_dA[(size_t)b * a.dim + d] = (dA2.x + dA2.y) * kLn2;
        a.part_dD[(size_t)b * a.dim + d] = dD2.x + dD2.y;
        a.part_dbias[(size_t)b * a.dim + d] = db2.x + db2.y;
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename T, int kG>
cudaError_t launch_bwd_wtma_g(const CUtensorMap *tm, const WinTmaArgs &a, int grid, bool dout_f32, cudaStream_t stream) {
    void (*kernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const WinTmaArgs);
    if (a.softplus) kernel = dout_f32 ? &ss_bwd_wtma_kernel<T, true, true, kG> : &ss_bwd_wtma_kernel<T, true, false, kG>;
    else kernel = dout_f32 ? &ss_bwd_wtma_kernel<T, false, true, kG> : &ss_bwd_wtma_kernel<T, false, false, kG>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
    if (e != cudaSuccess) return e;
    kernel<<<grid, 32, a.smem_bytes, stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], a);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_bwd_wtma(const CUtensorMap *tm, const WinTmaArgs &a, int grid, bool dout_f32, cudaStream_t stream) {
    return a.g == 1 ? launch_bwd_wtma_g<T, 1>(tm, a, grid, dout_f32, stream) : cudaErrorInvalidValue;
}

}  // namespace mia
