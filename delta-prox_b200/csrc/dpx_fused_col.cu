// dpx_fused_col.cu — column kernels of the fused FFT engine (k_col, k_col_tma, k_pack), one instantiation per column length.
#include "dpx_fused_launch.cuh"

namespace dpx {
namespace fused {
namespace launch {

template <class TH>
cudaError_t col(dim3 grid, size_t smem, const ColParams& p, cudaStream_t s) {
  cudaError_t e = prep(k_col<TH>, smem);
  if (e != cudaSuccess) return e;
  if (p.pdl) return launch_pdl(k_col<TH>, grid, ColThreads<TH>::value, smem, s, p);
  k_col<TH><<<grid, ColThreads<TH>::value, smem, s>>>(p);
  return cudaGetLastError();
}
template <class TH>
cudaError_t col_tma(dim3 grid, size_t smem, const ColParams& p, int n_tiles, int nb, cudaStream_t s) {
  cudaError_t e = prep(k_col_tma<TH>, smem);
  if (e != cudaSuccess) return e;
  k_col_tma<TH><<<grid, ColTmaCfg<TH>::NT, smem, s>>>(p, n_tiles, nb);
  return cudaGetLastError();
}
template <class TH, typename V>
cudaError_t pack(const V* src, V* dst, int planes, int H, int W, int G, V zero, cudaStream_t s) {
  const size_t total = (size_t)planes * (G + 1) * H * CG;
  k_pack<TH, V><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(src, dst, planes, H, W, G, zero);
  return cudaGetLastError();
}

#define DPX_INST_COL(N)                                                                                       \
  template cudaError_t col<typename TileFor<N, CG>::type>(dim3, size_t, const ColParams&, cudaStream_t);      \
  template cudaError_t pack<typename TileFor<N, CG>::type, float2>(const float2*, float2*, int, int, int, int, float2, cudaStream_t); \
  template cudaError_t pack<typename TileFor<N, CG>::type, float>(const float*, float*, int, int, int, int, float, cudaStream_t);
DPX_W_SIZES(DPX_INST_COL)
DPX_H_ONLY_SIZES(DPX_INST_COL)
#ifndef DPX_EXP_SIZES
template cudaError_t col_tma<typename TileFor<1024, CG>::type>(dim3, size_t, const ColParams&, int, int, cudaStream_t);
#endif
template cudaError_t col_tma<typename TileFor<2048, CG>::type>(dim3, size_t, const ColParams&, int, int, cudaStream_t);

}  // namespace launch
}  // namespace fused
}  // namespace dpx
