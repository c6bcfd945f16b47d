// fwd selective-scan kernels, f32 activations (one translation unit per dtype so they compile in parallel)
#include "scan_fwd_fast.cuh"
#include "scan_fwd_rows.cuh"
#include "scan_fwd_rowsn.cuh"
#include "scan_fwd_stream.cuh"
#include "scan_fwd_chunks.cuh"
#include "scan_fwd_cw.cuh"
namespace mia {
template cudaError_t launch_fwd_any<float>(const ScanArgs &, int, cudaStream_t);
template cudaError_t launch_fwd_rows<float>(const RowsArgs &, int, bool, cudaStream_t);
template cudaError_t launch_fwd_rowsn<float>(const RowsNArgs &, int, bool, cudaStream_t);
template cudaError_t launch_fwd_stream<float>(const StreamArgs &, int, bool, cudaStream_t);
template cudaError_t launch_fwd_chunks<float>(const ChunkArgs &, int, bool, cudaStream_t);
template cudaError_t launch_fwd_cw<float>(const CUtensorMap *, const CwFwdArgs &, int, bool, cudaStream_t);
}  // namespace mia
