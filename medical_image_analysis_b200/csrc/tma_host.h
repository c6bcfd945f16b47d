// Host-side helpers for TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no -lcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace mia {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn tma_encode_fn() {
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

// 2-D map of `rows` x `cols` elements (cols contiguous, row pitch `pitch_bytes`), box [box_rows][box_cols]; swizzle_bytes 128 / 64 / 32 / 0.
// esize 2 -> 16-bit element (bit pattern only: bf16 and fp16 move alike), 4 -> 32-bit.  Returns the CUresult (0 = success, -1 = no driver entry).
inline int tma_make_2d(CUtensorMap *tm, const void *ptr, unsigned long long rows, unsigned long long cols, unsigned long long pitch_bytes,
                       unsigned box_rows, unsigned box_cols, int esize, int swizzle_bytes = 128) {
    EncodeTiledFn fn = tma_encode_fn();
    if (!fn) return -1;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return (int)fn(tm, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void *>(ptr), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B :
                   swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace mia
