// dpx_fused_fft.cu — fused sm_100a FFT engine (placeholder until the kernels land).
#include "dpx_fft.cuh"

namespace dpx {

int make_fused_engine(const Geom& g, FftEngine** out) {
  (void)g;
  *out = nullptr;
  set_error("fused FFT engine: shape [%d x %d] not supported", g.H, g.W);
  return DPX_ERR_INVALID;
}

}  // namespace dpx
