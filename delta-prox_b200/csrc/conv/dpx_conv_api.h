// dpx_conv_api.h — plain declarations of the three tcgen05 implicit-GEMM convolutions of FFDNet-color
// (implemented in dpx_conv_{head,body,tail}.cu from conv/dpx_conv_umma.cuh).  NHWC bf16 activations, KRSC bf16 filters,
// fp32 bias; channels padded to 16 (head input, tail output) / 96.  Return 0 on success.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace dpx {
namespace conv {
int conv_head(const void* act, const void* flt, const float* bias, void* out, int n, int h, int w, void* ws, size_t wsb, cudaStream_t s);  // 16 -> 96, ReLU
int conv_body(const void* act, const void* flt, const float* bias, void* out, int n, int h, int w, void* ws, size_t wsb, cudaStream_t s);  // 96 -> 96, ReLU
int conv_tail(const void* act, const void* flt, const float* bias, void* out, int n, int h, int w, void* ws, size_t wsb, cudaStream_t s);  // 96 -> 16
size_t conv_workspace(int n, int h, int w);
}  // namespace conv
}  // namespace dpx
