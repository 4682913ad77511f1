// dpx_fused_rowzp.cu — the persistent row kernel of the plane-pair engine (k_rowz_mid_persist: PM_MID / PM_LAST / PM_XONLY / PM_FIRST), one
// instantiation per row length; split from dpx_fused_rowz.cu so that the two halves compile in parallel.
#include "dpx_fused_launch.cuh"

namespace dpx {
namespace fused {
namespace launch {

template <class TW, int PM>
cudaError_t rowz_persist(dim3 grid, size_t smem, const RowParams& p, int n_tiles, cudaStream_t s) {
  cudaError_t e = prep(k_rowz_mid_persist<TW, PM>, smem);
  if (e != cudaSuccess) return e;
  if (p.pdl) return launch_pdl(k_rowz_mid_persist<TW, PM>, grid, RowZPersistSmem<TW>::THREADS, smem, s, p, n_tiles);
  k_rowz_mid_persist<TW, PM><<<grid, RowZPersistSmem<TW>::THREADS, smem, s>>>(p, n_tiles);
  return cudaGetLastError();
}

#define DPX_INST_ROWP(N)                                                                                                   \
  template cudaError_t rowz_persist<typename TileFor<N, ZR>::type, PM_MID>(dim3, size_t, const RowParams&, int, cudaStream_t);   \
  template cudaError_t rowz_persist<typename TileFor<N, ZR>::type, PM_LAST>(dim3, size_t, const RowParams&, int, cudaStream_t);  \
  template cudaError_t rowz_persist<typename TileFor<N, ZR>::type, PM_XONLY>(dim3, size_t, const RowParams&, int, cudaStream_t); \
  template cudaError_t rowz_persist<typename TileFor<N, ZR>::type, PM_FIRST>(dim3, size_t, const RowParams&, int, cudaStream_t);
DPX_W_SIZES(DPX_INST_ROWP)

}  // namespace launch
}  // namespace fused
}  // namespace dpx
