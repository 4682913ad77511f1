#include "dpx_conv_umma.cuh"
#ifndef DPX_BODY_2SM
#define DPX_BODY_2SM true
#endif
namespace dpx { namespace conv {
int conv_body(const void* act, const void* flt, const float* bias, void* out, int n, int h, int w, void* ws, size_t wsb, cudaStream_t s) {
  return Conv3x3<96, 32, true, DPX_BODY_2SM>::run(act, flt, bias, out, n, h, w, 96, 96, ws, wsb, s);
}
size_t conv_workspace(int n, int h, int w) { return Conv3x3<96, 32, true, DPX_BODY_2SM>::workspace_size(n, h, w, 96, 96) + 1024; }
}}
