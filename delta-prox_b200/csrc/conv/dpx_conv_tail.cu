#include "dpx_conv_umma.cuh"
namespace dpx { namespace conv {
int conv_tail(const void* act, const void* flt, const float* bias, void* out, int n, int h, int w, void* ws, size_t wsb, cudaStream_t s) {
  return Conv3x3<16, 32, false>::run(act, flt, bias, out, n, h, w, 96, 16, ws, wsb, s);
}
}}
