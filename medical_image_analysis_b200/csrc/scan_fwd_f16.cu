// fwd selective-scan kernels, f16 activations (one translation unit per dtype so they compile in parallel)
#include "scan_fwd_fast.cuh"
#include "scan_fwd_rows.cuh"
#include "scan_fwd_rowsn.cuh"
#include "scan_fwd_stream.cuh"
#include "scan_fwd_chunks.cuh"
#include "scan_fwd_cw.cuh"
namespace mia {
template cudaError_t launch_fwd_any<__half>(const ScanArgs &, int, cudaStream_t);
template cudaError_t launch_fwd_rows<__half>(const RowsArgs &, int, bool, cudaStream_t);
template cudaError_t launch_fwd_rowsn<__half>(const RowsNArgs &, int, bool, cudaStream_t);
template cudaError_t launch_fwd_stream<__half>(const StreamArgs &, int, bool, cudaStream_t);
template cudaError_t launch_fwd_chunks<__half>(const ChunkArgs &, int, bool, cudaStream_t);
template cudaError_t launch_fwd_cw<__half>(const CUtensorMap *, const CwFwdArgs &, int, bool, cudaStream_t);
}  // namespace mia
