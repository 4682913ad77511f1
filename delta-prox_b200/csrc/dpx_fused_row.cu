// dpx_fused_row.cu — row kernels of the half-spectrum engine (k_row, k_row_mid_persist), one instantiation per row length.
#include "dpx_fused_launch.cuh"

namespace dpx {
namespace fused {
namespace launch {

template <class TW, int MODE, bool SINGLE>
cudaError_t row(dim3 grid, size_t smem, const RowParams& p, cudaStream_t s) {
  cudaError_t e = prep(k_row<TW, MODE, SINGLE>, smem);
  if (e != cudaSuccess) return e;
  if (p.pdl) return launch_pdl(k_row<TW, MODE, SINGLE>, grid, kThreads, smem, s, p);
  k_row<TW, MODE, SINGLE><<<grid, kThreads, smem, s>>>(p);
  return cudaGetLastError();
}
template <class TW>
cudaError_t row_persist(dim3 grid, size_t smem, const RowParams& p, int n_tiles, cudaStream_t s) {
  cudaError_t e = prep(k_row_mid_persist<TW>, smem);
  if (e != cudaSuccess) return e;
  if (p.pdl) return launch_pdl(k_row_mid_persist<TW>, grid, kThreads, smem, s, p, n_tiles);
  k_row_mid_persist<TW><<<grid, kThreads, smem, s>>>(p, n_tiles);
  return cudaGetLastError();
}

// the combinations the driver uses: FIRST and XONLY only in their general (accumulating) form, MID / LAST in both
#define DPX_INST_ROW(N)                                                                                                    \
  template cudaError_t row<typename TileFor<N, ROWS / 2>::type, ROW_FIRST, false>(dim3, size_t, const RowParams&, cudaStream_t); \
  template cudaError_t row<typename TileFor<N, ROWS / 2>::type, ROW_MID, true>(dim3, size_t, const RowParams&, cudaStream_t);    \
  template cudaError_t row<typename TileFor<N, ROWS / 2>::type, ROW_MID, false>(dim3, size_t, const RowParams&, cudaStream_t);   \
  template cudaError_t row<typename TileFor<N, ROWS / 2>::type, ROW_LAST, true>(dim3, size_t, const RowParams&, cudaStream_t);   \
  template cudaError_t row<typename TileFor<N, ROWS / 2>::type, ROW_LAST, false>(dim3, size_t, const RowParams&, cudaStream_t);  \
  template cudaError_t row<typename TileFor<N, ROWS / 2>::type, ROW_XONLY, false>(dim3, size_t, const RowParams&, cudaStream_t); \
  template cudaError_t row_persist<typename TileFor<N, ROWS / 2>::type>(dim3, size_t, const RowParams&, int, cudaStream_t);
DPX_W_SIZES(DPX_INST_ROW)

}  // namespace launch
}  // namespace fused
}  // namespace dpx
