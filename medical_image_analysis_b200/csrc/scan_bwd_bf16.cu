// bwd selective-scan kernels, bf16 activations (one translation unit per dtype so they compile in parallel)
#include "scan_bwd_fast.cuh"
#include "scan_bwd_rows.cuh"
#include "scan_bwd_rowsn.cuh"
#include "scan_bwd_cw.cuh"
namespace mia {
template cudaError_t launch_bwd_any<__nv_bfloat16>(const ScanArgs &, int, cudaStream_t);
template cudaError_t launch_bwd_rows<__nv_bfloat16>(const RowsBwdArgs &, int, bool, cudaStream_t);
template cudaError_t launch_bwd_rowsn<__nv_bfloat16>(const RowsNBwdArgs &, int, bool, cudaStream_t);
template cudaError_t launch_bwd_cw<__nv_bfloat16>(const CUtensorMap *, const CwBwdArgs &, int, bool, cudaStream_t);
}  // namespace mia
