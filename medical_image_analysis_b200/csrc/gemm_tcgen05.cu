// Tensor-core GEMM for the projections around the scan and for the ViT-MAE patch encode (sm_100a: tcgen05 + TMEM + TMA).
//
//   C[M, N] = act( A[M, K] . W[N, K]^T + bias[N] )        A, W: bf16 / fp16, K contiguous ("TN", nn.Linear's layout)
//
// Replaces the cuBLAS / cuDNN calls behind nn.Linear / einsum / kernel==stride nn.Conv2d at:
//   SS2D in_proj / out_proj / x_proj            R2GenCSR/VMamba/classification/models/vmamba.py:751, 775, 386
//   Mamba in_proj / x_proj / dt_proj / out_proj */arm/Finetuning/mamba_simple.py:408-414, 686-689, 708
//   SmallPatchEmbed conv16/16, conv4/4, conv1x1 HD_Xray_Pretrain_MAE/pretrain/patch_embed.py:25-41 (as GEMMs over patches)
//   timm Block qkv / proj / fc1 / fc2           HD_Xray_Pretrain_MAE/pretrain/models/mae.py:64-66, 82-84
//
// Structure (one persistent CTA per SM, 192 threads, warp-specialised; Blackwell guide "anatomy"):
//   warp 0   TMA producer: 2-D tensor-map loads (cp.async.bulk.tensor.2d, SASS UTMALDG) of [128 x 64] A and [BN x 64] W
//            tiles with the 128-byte swizzle into a ring of kStages shared-memory stages (full / empty mbarriers);
//   warp 1   allocates TMEM (2 accumulator stages of BN fp32 columns) and issues tcgen05.mma.cta_group::1.kind::f16
//            (SASS UTCHMMA), M = 128, N = BN, K = 16 per instruction, four per 64-wide K block; tcgen05.commit releases the
//            shared-memory stage and, after the last K block, hands the accumulator stage to the epilogue;
//   warps 2-5  epilogue: tcgen05.ld (SASS LDTM) 32 lanes x 32 columns at a time, + bias, activation, conversion, then
//            128-byte row chunks into a swizzled shared slab and out through a TMA tensor-map store (SASS UTMASTG:
//            coalesced full-line writes; per-thread 16-byte global stores only when C's rows are not 16-byte aligned);
//            the next tile's MMAs run meanwhile in the other accumulator stage.
// Out-of-range rows / columns / K tail are zero-filled by TMA and masked in the epilogue, so any M, N and any K with
// K % 8 == 0 (16-byte rows) work.  No split-K: K is at most 16384 (patch-embed conv4/4) and M is large on this path.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/mia_gemm.h"
#include "scan_common.cuh"

namespace {

constexpr int kBM = 128, kBK = 64;
constexpr int kGemmThreads = 192;

thread_local char g_gemm_err[384] = "";
int gfail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_gemm_err, sizeof(g_gemm_err), fmt, ap);
    va_end(ap);
    return code;
}

struct GemmArgs {
    int M, N, K;
    int act, out_f32, has_bias, in_f16;
    const float *bias;
    void *C;
    long long ldc;
    int num_m_blocks, num_n_blocks, num_k_blocks;
    int tma_store;      // C is written through a tensor map (16-byte aligned rows); else direct 16-byte stores per thread
};

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     mia::smem_u32(smem_dst)),
                 "l"(tm), "r"(mia::smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, const void *smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(mia::smem_u32(smem_src)), "r"(c0),
                 "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mia::smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T ; accumulate == 0 overwrites
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major tile staged by TMA with the 128-byte swizzle: rows of 64 elements
// (128 bytes), 8-row groups 1024 bytes apart (SBO), version 1 (Blackwell), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

// MN-major operand (the M / N index is the contiguous one, e.g. X^T or W^T read in place): TMA boxes of [64 k-rows x 64
// mn-elements (128 bytes)] with the 128-byte swizzle, consecutive 64-wide mn slices 8192 bytes apart.  Canonical layout
// ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) elements: LBO = 8192 B between mn slices, SBO = 1024 B between groups of 8 k-rows.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

__device__ __forceinline__ float act_apply(float v, int act) {
    switch (act) {
        case MIA_ACT_RELU: return fmaxf(v, 0.f);
        case MIA_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
        case MIA_ACT_SILU: return v / (1.f + __expf(-v));
        default: return v;
    }
}

constexpr int kEpiSlab = 32 * 128;       // one epilogue staging slab: 32 rows x 128 bytes (one swizzle-128B TMA store box)

template <int BN, int kStages, bool kAMN, bool kBMN>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmC, const GemmArgs g) {
    extern __shared__ char smem_raw[];
    // SWIZZLE_128B tiles must sit on 1024-byte boundaries
    char *smem = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int kABytes = kBM * kBK * 2, kBBytes = BN * kBK * 2, kStageBytes = kABytes + kBBytes;
    char *tiles = smem;
    char *epi = smem + kStages * kStageBytes;                 // 4 warps x 2 slabs, 1024-byte aligned (swizzle-128B boxes)
    uint64_t *full = reinterpret_cast<uint64_t *>(epi + 8 * kEpiSlab);
    uint64_t *empty = full + kStages;
    uint64_t *tfull = empty + kStages;        // accumulator stage ready for the epilogue
    uint64_t *tempty = tfull + 2;             // accumulator stage drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mia::mbar_init(full + s, 1); mia::mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mia::mbar_init(tfull + s, 1); mia::mbar_init(tempty + s, 4); }
        mia::fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mia::smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_tiles = g.num_m_blocks * g.num_n_blocks;
    const int KB = g.num_k_blocks;

    if (warp == 0) {
        // ===== TMA producer
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mb = tile % g.num_m_blocks, nb = tile / g.num_m_blocks;     // m fastest: a column of W stays in L2
                for (int kb = 0; kb < KB; ++kb) {
                    mia::mbar_wait(empty + s, ph ^ 1);
                    mia::mbar_arrive_expect_tx(full + s, kStageBytes);
                    char *sa = tiles + s * kStageBytes, *sb = sa + kABytes;
                    if (kAMN) {
#pragma unroll
                        for (int i = 0; i < kBM / 64; ++i) tma_load_2d(sa + i * 8192, &tmA, mb * kBM + 64 * i, kb * kBK, full + s);
                    } else {
                        tma_load_2d(sa, &tmA, kb * kBK, mb * kBM, full + s);
                    }
                    if (kBMN) {
#pragma unroll
                        for (int i = 0; i < BN / 64; ++i) tma_load_2d(sb + i * 8192, &tmB, nb * BN + 64 * i, kb * kBK, full + s);
                    } else {
                        tma_load_2d(sb, &tmB, kb * kBK, nb * BN, full + s);
                    }
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected lane)
        if (lane == 0) {
            // instruction descriptor: D fp32, A / B bf16 (or fp16), both K-major, N = BN, M = 128
            const uint32_t fmt = g.in_f16 ? 0u : 1u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((kAMN ? 1u : 0u) << 15) | ((kBMN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
            int s = 0, as = 0;
            uint32_t ph = 0, aph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mia::mbar_wait(tempty + as, aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < KB; ++kb) {
                    mia::mbar_wait(full + s, ph);
                    tc_fence_after();
                    const uint32_t a_addr = mia::smem_u32(tiles + s * kStageBytes), b_addr = a_addr + kABytes;
                    const uint64_t da = kAMN ? umma_desc_mn_sw128(a_addr) : umma_desc_k_sw128(a_addr);
                    const uint64_t db = kBMN ? umma_desc_mn_sw128(b_addr) : umma_desc_k_sw128(b_addr);
                    // 16 elements along K per instruction: K-major = 32 bytes inside the swizzle atom (+2 in the address field);
                    // MN-major = 16 k-rows of 128 bytes = 2048 bytes (+128)
                    constexpr uint64_t ka = kAMN ? 128 : 2, kbb = kBMN ? 128 : 2;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)
                        tc_mma_f16(tmem_d, da + ka * k, db + kbb * k, idesc, (kb | k) != 0);
                    tc_commit(empty + s);                   // the stage is free once these MMAs have read it
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
                tc_commit(tfull + as);                      // accumulator complete
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lanes of quadrant (warp % 4)
        const int q = warp & 3;
        int as = 0, chunk = 0;
        uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int mb = tile % g.num_m_blocks, nb = tile / g.num_m_blocks;
            mia::mbar_wait(tfull + as, aph);
            tc_fence_after();
            const int row = mb * kBM + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
            if (g.tma_store) {
                // 128-byte chunks of the 32-row slab of this warp: registers -> swizzled shared slab -> one TMA store each.
                // Chunk j of row r sits at r * 128 + ((j ^ (r & 7)) << 4): conflict-free 16-byte stores, and exactly the
                // SWIZZLE_128B pattern the C tensor map un-does on the way out (coalesced full-line writes).
                const int cc = g.out_f32 ? 32 : 64;                        // columns per chunk
                char *slab0 = epi + (warp - 2) * 2 * kEpiSlab;
                for (int c = 0; c < BN / cc; ++c, ++chunk) {
                    const int col0 = nb * BN + c * cc;
                    if (col0 >= g.N) break;                                 // warp-uniform
                    char *slab = slab0 + (chunk & 1) * kEpiSlab;
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that used this slab 2 chunks ago
                    __syncwarp();
                    char *rowp = slab + lane * 128;
#pragma unroll 1
                    for (int hlf = 0; hlf < (g.out_f32 ? 1 : 2); ++hlf) {
                        uint32_t r[32];
                        tc_ld32(taddr + (uint32_t)(c * cc + hlf * 32), r);
                        tc_wait_ld();
                        float v[32];
                        const int cb = col0 + hlf * 32;
                        if (g.has_bias) {
                            if (cb + 32 <= g.N && ((reinterpret_cast<uintptr_t>(g.bias + cb) & 15) == 0)) {   // 8 x 16-byte loads
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float4 b4 = __ldg(reinterpret_cast<const float4 *>(g.bias + cb) + j);
                                    v[4 * j] = __uint_as_float(r[4 * j]) + b4.x; v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
                                    v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z; v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + ((cb + j < g.N) ? __ldg(g.bias + cb + j) : 0.f);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                        }
                        if (g.act == MIA_ACT_RELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                        } else if (g.act != MIA_ACT_NONE) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = act_apply(v[j], g.act);
                        }
                        if (g.out_f32) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                *reinterpret_cast<float4 *>(rowp + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        } else {
                            uint32_t w[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                if (g.in_f16) { __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]); w[j] = *reinterpret_cast<uint32_t *>(&h); }
                                else { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]); w[j] = *reinterpret_cast<uint32_t *>(&h); }
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                *reinterpret_cast<uint4 *>(rowp + (((hlf * 4 + j) ^ (lane & 7)) << 4)) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                        }
                    }
                    mia::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmC, slab, col0, mb * kBM + q * 32);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            } else {
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                tc_ld32(taddr + (uint32_t)(c * 32), r);
                tc_wait_ld();
                const int col0 = nb * BN + c * 32;
                if (row < g.M && col0 < g.N) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        v[j] = __uint_as_float(r[j]);
                        if (g.has_bias) v[j] += (col0 + j < g.N) ? __ldg(g.bias + col0 + j) : 0.f;
                        v[j] = act_apply(v[j], g.act);
                    }
                    const int nvalid = min(32, g.N - col0);
                    if (g.out_f32) {
                        float *dst = reinterpret_cast<float *>(g.C) + (size_t)row * g.ldc + col0;
                        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) reinterpret_cast<float4 *>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        } else {
                            for (int j = 0; j < nvalid; ++j) dst[j] = v[j];
                        }
                    } else {
                        uint16_t *dst = reinterpret_cast<uint16_t *>(g.C) + (size_t)row * g.ldc + col0;
                        uint32_t w[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (g.in_f16) { __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]); w[j] = *reinterpret_cast<uint32_t *>(&h); }
                            else { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]); w[j] = *reinterpret_cast<uint32_t *>(&h); }
                        }
                        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) reinterpret_cast<uint4 *>(dst)[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                        } else {
                            for (int j = 0; j < nvalid; ++j) dst[j] = (uint16_t)((w[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
                        }
                    }
                }
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mia::mbar_arrive(tempty + as);
            if (++as == 2) { as = 0; aph ^= 1; }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");    // the slabs must outlive their stores
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D map of a row-major [rows][cols] matrix with row pitch ld (elements); box = [box_rows][128 bytes of columns], 128B swizzle.
// dt: MIA_GEMM_F32 / F16 / BF16
int make_map(CUtensorMap *tm, const void *ptr, long long rows, long long cols, long long ld, int box_rows, int dt) {
    // rows x cols with cols contiguous; the box is [box_rows][128 bytes of columns]
    EncodeTiledFn fn = encode_fn();
    if (!fn) return gfail(MIA_GEMM_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const int es = dt == MIA_GEMM_F32 ? 4 : 2;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType cdt = dt == MIA_GEMM_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                        : (dt == MIA_GEMM_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
    const CUresult r = fn(tm, cdt, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return gfail(MIA_GEMM_ECUDA, "cuTensorMapEncodeTiled failed (CUresult %d): rows %lld cols %lld ld %lld", (int)r, rows, cols, ld);
    return MIA_GEMM_OK;
}

template <int BN, int kStages, bool kAMN, bool kBMN>
int launch_gemm(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, GemmArgs &g, int sms, cudaStream_t stream) {
    constexpr int smem = kStages * (kBM * kBK * 2 + BN * kBK * 2) + 8 * kEpiSlab + (2 * kStages + 4) * 8 + 16 + 1024;
    auto k = &gemm_tn_kernel<BN, kStages, kAMN, kBMN>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return gfail(MIA_GEMM_ECUDA, "cudaFuncSetAttribute(gemm, %d B): %s", smem, cudaGetErrorString(cudaGetLastError()));
    g.num_m_blocks = (g.M + kBM - 1) / kBM;
    g.num_n_blocks = (g.N + BN - 1) / BN;
    g.num_k_blocks = (g.K + kBK - 1) / kBK;
    const int tiles = g.num_m_blocks * g.num_n_blocks;
    const int grid = tiles < sms ? tiles : sms;
    k<<<grid, kGemmThreads, smem, stream>>>(tmA, tmB, tmC, g);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return gfail(MIA_GEMM_ECUDA, "gemm launch: %s", cudaGetErrorString(e));
    return MIA_GEMM_OK;
}

}  // namespace

static int gemm_impl(const void *A, const void *W, const float *bias, void *C, int M, int N, int K, long long lda, long long ldw, long long ldc,
                     int a_mn, int b_mn, int in_dtype, int out_dtype, int act, void *cuda_stream) {
    if (!A || !W || !C) return gfail(MIA_GEMM_EINVAL, "gemm: null pointer");
    if (M <= 0 || N <= 0 || K <= 0) return gfail(MIA_GEMM_EINVAL, "gemm: empty or negative size (M %d, N %d, K %d)", M, N, K);
    if (in_dtype != MIA_GEMM_BF16 && in_dtype != MIA_GEMM_F16) return gfail(MIA_GEMM_EINVAL, "gemm: inputs must be bf16 or fp16");
    if (out_dtype != in_dtype && out_dtype != MIA_GEMM_F32) return gfail(MIA_GEMM_EINVAL, "gemm: output must be the input dtype or fp32");
    if (act < 0 || act > MIA_ACT_SILU) return gfail(MIA_GEMM_EINVAL, "gemm: unknown activation %d", act);
    if ((lda % 8) || (ldw % 8) || lda < (a_mn ? M : K) || ldw < (b_mn ? N : K))
        return gfail(MIA_GEMM_EINVAL, "gemm: row pitches must cover a row and be multiples of 8 elements (lda %lld, ldw %lld, M %d, N %d, K %d)", lda, ldw, M, N, K);
    if (((uintptr_t)A | (uintptr_t)W) & 15) return gfail(MIA_GEMM_EINVAL, "gemm: A and W must be 16-byte aligned");
    if (ldc < N) return gfail(MIA_GEMM_EINVAL, "gemm: ldc < N");
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return gfail(MIA_GEMM_ECUDA, "gemm: cannot query the device");
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.M = M; g.N = N; g.K = K; g.act = act; g.out_f32 = out_dtype == MIA_GEMM_F32; g.has_bias = bias != nullptr; g.bias = bias;
    g.in_f16 = in_dtype == MIA_GEMM_F16; g.C = C; g.ldc = ldc;
    const int BN = N <= 64 ? 64 : (N <= 128 || (long long)((M + 127) / 128) * ((N + 255) / 256) < sms ? 128 : 256);
    CUtensorMap tmA, tmB, tmC;
    // K-major operand: [rows = M or N][cols = K], box [tile rows][64 k];  MN-major operand: stored [rows = K][cols = M or N], box [64 k][64 mn]
    if (int rc = a_mn ? make_map(&tmA, A, K, M, lda, kBK, in_dtype) : make_map(&tmA, A, M, K, lda, kBM, in_dtype)) return rc;
    if (int rc = b_mn ? make_map(&tmB, W, K, N, ldw, kBK, in_dtype) : make_map(&tmB, W, N, K, ldw, BN, in_dtype)) return rc;
    const int eo = g.out_f32 ? 4 : 2;
    g.tma_store = (((uintptr_t)C & 15) == 0) && ((ldc * eo) % 16 == 0);
    if (g.tma_store) {
        if (int rc = make_map(&tmC, C, M, N, ldc, 32, out_dtype)) return rc;
    } else {
        tmC = tmA;                                   // unused by the kernel on the direct-store path
    }
    cudaStream_t stream = (cudaStream_t)cuda_stream;
#define MIA_GEMM_LAUNCH(AMN, BMN)                                                          \
    switch (BN) {                                                                          \
        case 64: return launch_gemm<64, 8, AMN, BMN>(tmA, tmB, tmC, g, sms, stream);       \
        case 128: return launch_gemm<128, 6, AMN, BMN>(tmA, tmB, tmC, g, sms, stream);     \
        default: return launch_gemm<256, 4, AMN, BMN>(tmA, tmB, tmC, g, sms, stream);      \
    }
    if (!a_mn && !b_mn) { MIA_GEMM_LAUNCH(false, false) }
    if (!a_mn && b_mn) { MIA_GEMM_LAUNCH(false, true) }
    if (a_mn && b_mn) { MIA_GEMM_LAUNCH(true, true) }
    MIA_GEMM_LAUNCH(true, false)
#undef MIA_GEMM_LAUNCH
}

extern "C" {

const char *mia_gemm_last_error(void) { return g_gemm_err; }

int mia_gemm_tn(const void *A, const void *W, const float *bias, void *C, int M, int N, int K, long long lda, long long ldw, long long ldc,
                int in_dtype, int out_dtype, int act, void *cuda_stream) {
    return gemm_impl(A, W, bias, C, M, N, K, lda, ldw, ldc, 0, 0, in_dtype, out_dtype, act, cuda_stream);
}

int mia_gemm(const void *A, const void *B, const float *bias, void *C, int M, int N, int K, long long lda, long long ldb, long long ldc,
             int a_mn_major, int b_mn_major, int in_dtype, int out_dtype, int act, void *cuda_stream) {
    return gemm_impl(A, B, bias, C, M, N, K, lda, ldb, ldc, a_mn_major != 0, b_mn_major != 0, in_dtype, out_dtype, act, cuda_stream);
}

}  // extern "C"
