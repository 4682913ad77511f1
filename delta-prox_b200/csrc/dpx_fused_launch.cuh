// dpx_fused_launch.cuh — launchers of the fused-engine kernels, declared here and explicitly instantiated for every supported size in
// four translation units (dpx_fused_col.cu, dpx_fused_row.cu, dpx_fused_rowz.cu, dpx_fused_rowzp.cu) so that they compile in parallel; the engine
// (dpx_fused_fft.cu) only sees the declarations.  Each launcher sets the dynamic shared-memory attribute, launches on `s` and
// returns the launch status.
#pragma once
#include <cuda_runtime.h>

#include "dpx_fused_driver.cuh"

namespace dpx {
namespace fused {
namespace launch {

template <class TW, int MODE, bool SINGLE> cudaError_t row(dim3 grid, size_t smem, const RowParams& p, cudaStream_t s);
template <class TW> cudaError_t row_persist(dim3 grid, size_t smem, const RowParams& p, int n_tiles, cudaStream_t s);
template <class TW, int MODE, bool SINGLE> cudaError_t rowz(dim3 grid, size_t smem, const RowParams& p, cudaStream_t s);
template <class TW, int PM> cudaError_t rowz_persist(dim3 grid, size_t smem, const RowParams& p, int n_tiles, cudaStream_t s);
template <class TH> cudaError_t col(dim3 grid, size_t smem, const ColParams& p, cudaStream_t s);
template <class TH> cudaError_t col_tma(dim3 grid, size_t smem, const ColParams& p, int n_tiles, int nb, cudaStream_t s);
template <class TH, typename V> cudaError_t pack(const V* src, V* dst, int planes, int H, int W, int G, V zero, cudaStream_t s);

// launch as a programmatic dependent of the previous kernel in the stream (the kernel calls griddepcontrol.wait: see griddep_wait)
template <class K, class... A>
inline cudaError_t launch_pdl(K kernel, dim3 grid, int threads, size_t smem, cudaStream_t s, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// shared by the instantiating translation units
template <class K>
inline cudaError_t prep(K kernel, size_t smem) {
  return smem > 48 * 1024 ? cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) : cudaSuccess;
}

}  // namespace launch
}  // namespace fused
}  // namespace dpx
