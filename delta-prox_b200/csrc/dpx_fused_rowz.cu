// dpx_fused_rowz.cu — the non-persistent row kernel of the plane-pair engine (k_rowz), one instantiation per row length
// (k_rowz_mid_persist: dpx_fused_rowzp.cu).
#include "dpx_fused_launch.cuh"

namespace dpx {
namespace fused {
namespace launch {

template <class TW, int MODE, bool SINGLE>
cudaError_t rowz(dim3 grid, size_t smem, const RowParams& p, cudaStream_t s) {
  cudaError_t e = prep(k_rowz<TW, MODE, SINGLE>, smem);
  if (e != cudaSuccess) return e;
  if (p.pdl) return launch_pdl(k_rowz<TW, MODE, SINGLE>, grid, kThreads, smem, s, p);
  k_rowz<TW, MODE, SINGLE><<<grid, kThreads, smem, s>>>(p);
  return cudaGetLastError();
}

// the combinations the driver uses: FIRST and XONLY only in their general (accumulating) form, MID / LAST in both
#define DPX_INST_ROW(N)                                                                                                    \
  template cudaError_t rowz<typename TileFor<N, ROWS>::type, ROW_FIRST, false>(dim3, size_t, const RowParams&, cudaStream_t); \
  template cudaError_t rowz<typename TileFor<N, ROWS>::type, ROW_MID, true>(dim3, size_t, const RowParams&, cudaStream_t);    \
  template cudaError_t rowz<typename TileFor<N, ROWS>::type, ROW_MID, false>(dim3, size_t, const RowParams&, cudaStream_t);   \
  template cudaError_t rowz<typename TileFor<N, ROWS>::type, ROW_LAST, true>(dim3, size_t, const RowParams&, cudaStream_t);   \
  template cudaError_t rowz<typename TileFor<N, ROWS>::type, ROW_LAST, false>(dim3, size_t, const RowParams&, cudaStream_t);  \
  template cudaError_t rowz<typename TileFor<N, ROWS>::type, ROW_XONLY, false>(dim3, size_t, const RowParams&, cudaStream_t);
DPX_W_SIZES(DPX_INST_ROW)

}  // namespace launch
}  // namespace fused
}  // namespace dpx
