// dpx_fused_rowz.cu — row kernels of the plane-pair engine (k_rowz, k_rowz_mid_persist), one instantiation per row length.
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
template <class TW, int PM>
cudaError_t rowz_persist(dim3 grid, size_t smem, const RowParams& p, int n_tiles, cudaStream_t s) {
  cudaError_t e = prep(k_rowz_mid_persist<TW, PM>, smem);
  if (e != cudaSuccess) return e;
  if (p.pdl) return launch_pdl(k_rowz_mid_persist<TW, PM>, grid, RowZPersistSmem<TW>::THREADS, smem, s, p, n_tiles);
  k_rowz_mid_persist<TW, PM><<<grid, RowZPersistSmem<TW>::THREADS, smem, s>>>(p, n_tiles);
  return cudaGetLastError();
}

// the combinations the driver uses: FIRST and XONLY only in their general (accumulating) form, MID / LAST in both
#define DPX_INST_ROW(N)                                                                                                    \
  template cudaError_t rowz<typename TileFor<N, ROWS>::type, ROW_FIRST, false>(dim3, size_t, const RowParams&, cudaStream_t); \
  template cudaError_t rowz<typename TileFor<N, ROWS>::type, ROW_MID, true>(dim3, size_t, const RowParams&, cudaStream_t);    \
  template cudaError_t rowz<typename TileFor<N, ROWS>::type, ROW_MID, false>(dim3, size_t, const RowParams&, cudaStream_t);   \
  template cudaError_t rowz<typename TileFor<N, ROWS>::type, ROW_LAST, true>(dim3, size_t, const RowParams&, cudaStream_t);   \
  template cudaError_t rowz<typename TileFor<N, ROWS>::type, ROW_LAST, false>(dim3, size_t, const RowParams&, cudaStream_t);  \
  template cudaError_t rowz<typename TileFor<N, ROWS>::type, ROW_XONLY, false>(dim3, size_t, const RowParams&, cudaStream_t); \
  template cudaError_t rowz_persist<typename TileFor<N, ZR>::type, PM_MID>(dim3, size_t, const RowParams&, int, cudaStream_t);   \
  template cudaError_t rowz_persist<typename TileFor<N, ZR>::type, PM_LAST>(dim3, size_t, const RowParams&, int, cudaStream_t);  \
  template cudaError_t rowz_persist<typename TileFor<N, ZR>::type, PM_XONLY>(dim3, size_t, const RowParams&, int, cudaStream_t); \
  template cudaError_t rowz_persist<typename TileFor<N, ZR>::type, PM_FIRST>(dim3, size_t, const RowParams&, int, cudaStream_t);
DPX_W_SIZES(DPX_INST_ROW)

}  // namespace launch
}  // namespace fused
}  // namespace dpx
