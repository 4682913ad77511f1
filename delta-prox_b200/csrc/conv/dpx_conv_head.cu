#include "dpx_conv_umma.cuh"
namespace dpx { namespace conv {
int conv_head(const void* act, const void* flt, const float* bias, void* out, int n, int h, int w, void* ws, size_t wsb, cudaStream_t s) {
  return Conv3x3<96, 16, true>::run(act, flt, bias, out, n, h, w, 16, 96, ws, wsb, s);
}
}}
