// fwd selective-scan kernels, f32 activations (one translation unit per dtype so they compile in parallel)
#include "scan_fwd_fast.cuh"
namespace mia {
template cudaError_t launch_fwd_any<float>(const ScanArgs &, int, cudaStream_t);
}  // namespace mia
