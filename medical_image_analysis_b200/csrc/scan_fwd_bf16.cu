// fwd selective-scan kernels, bf16 activations (one translation unit per dtype so they compile in parallel)
#include "scan_fwd_fast.cuh"
#include "scan_fwd_rows.cuh"
#include "scan_fwd_rowsn.cuh"
#include "scan_fwd_stream.cuh"
#include "scan_fwd_chunks.cuh"
#include "scan_fwd_cw.cuh"
namespace mia {
template cudaError_t launch_fwd_any<__nv_bfloat16>(const ScanArgs &, int, cudaStream_t);
template cudaError_t launch_fwd_rows<__nv_bfloat16>(const RowsArgs &, int, bool, cudaStream_t);
template cudaError_t launch_fwd_rowsn<__nv_bfloat16>(const RowsNArgs &, int, bool, cudaStream_t);
template cudaError_t launch_fwd_stream<__nv_bfloat16>(const StreamArgs &, int, bool, cudaStream_t);
template cudaError_t launch_fwd_chunks<__nv_bfloat16>(const ChunkArgs &, int, bool, cudaStream_t);
template cudaError_t launch_fwd_cw<__nv_bfloat16>(const CUtensorMap *, const CwFwdArgs &, int, bool, cudaStream_t);
}  // namespace mia
