// dpx_ffdnet.cu — native FFDNet-color forward (the denoiser behind deep_prior):
//   PixelUnshuffle(2) + sigma map -> conv3x3(13->96)+ReLU -> 10 x [conv3x3(96->96)+ReLU] -> conv3x3(96->12) -> PixelShuffle(2)
//   (proxfn/pnp/denoisers/models/network_ffdnet.py:44-68)
// Activations live in HBM as NHWC bf16 (channels padded to 16 / 96), the twelve convolutions run as implicit GEMMs on
// tcgen05 tensor cores with fp32 accumulation in TMEM (conv/dpx_conv_umma.cuh); the unshuffle/sigma prologue and the
// shuffle/crop epilogue are fused layout kernels.  bf16 operands => ~1e-2 relative accuracy: this is the opt-in FAST
// denoiser; the fp32 parity path keeps the framework convolution (dprox_b200/denoisers.py).
#include <cuda_bf16.h>

#include <new>

#include "dpx_common.cuh"
#ifndef DPX_NO_UMMA_CONV
#include "conv/dpx_conv_api.h"
#endif

using namespace dpx;

struct dpx_ffdnet {
  int nb = 0, nc = 0;
  __nv_bfloat16* w[32] = {nullptr};      // KRSC, channels padded
  float* bias[32] = {nullptr};
  bool set[32] = {false};
  __nv_bfloat16 *act0 = nullptr, *act1 = nullptr, *io16 = nullptr, *out16 = nullptr;
  void* ws = nullptr;
  size_t ws_bytes = 0, act_cap = 0, io_cap = 0;
};

namespace {

constexpr int CIN_PAD = 16, COUT_TAIL_PAD = 16;

// [Cout,Cin,3,3] fp32 -> [Cout_pad,3,3,Cin_pad] bf16 (zero padded)
__global__ void k_pack_filter(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cin, int cout_pad,
                              int cin_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = cout_pad * 9 * cin_pad;
  if (i >= total) return;
  const int c = i % cin_pad, rs = (i / cin_pad) % 9, k = i / (cin_pad * 9);
  const float v = (k < cout && c < cin) ? w[((size_t)k * cin + c) * 9 + rs] : 0.f;
  out[i] = __float2bfloat16(v);
}

__global__ void k_pad_bias(const float* __restrict__ b, float* __restrict__ out, int cout, int cout_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cout_pad) out[i] = i < cout ? b[i] : 0.f;
}

// x [B,3,H,W] fp32 -> NHWC bf16 [B,h2,w2,16]: channel c*4+dy*2+dx = x[b,c,2h+dy,2w+dx] (replicate-padded to even sizes),
// channel 12 = sigma_b, channels 13..15 = 0                                         network_ffdnet.py:54-64
__global__ void k_unshuffle_in(const float* __restrict__ x, const float* __restrict__ sigma, int sigma_per_sample,
                               __nv_bfloat16* __restrict__ out, int B, int H, int W, int h2, int w2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;          // one thread per output pixel
  const size_t total = (size_t)B * h2 * w2;
  if (i >= total) return;
  const int w = (int)(i % w2), h = (int)((i / w2) % h2), b = (int)(i / ((size_t)w2 * h2));
  __align__(16) __nv_bfloat16 v[16];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yy = min(2 * h + dy, H - 1), xx = min(2 * w + dx, W - 1);
        v[c * 4 + dy * 2 + dx] = __float2bfloat16(x[(((size_t)b * 3 + c) * H + yy) * W + xx]);
      }
  v[12] = __float2bfloat16(sigma[sigma_per_sample ? b : 0]);
  v[13] = v[14] = v[15] = __float2bfloat16(0.f);
  uint4* dst = reinterpret_cast<uint4*>(out + i * 16);
  dst[0] = reinterpret_cast<const uint4*>(v)[0];
  dst[1] = reinterpret_cast<const uint4*>(v)[1];
}

// NHWC bf16 [B,h2,w2,16] -> y [B,3,H,W] fp32 (PixelShuffle(2) + crop)              network_ffdnet.py:65-68
__global__ void k_shuffle_out(const __nv_bfloat16* __restrict__ in, float* __restrict__ y, int B, int H, int W, int h2, int w2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * h2 * w2;
  if (i >= total) return;
  const int w = (int)(i % w2), h = (int)((i / w2) % h2), b = (int)(i / ((size_t)w2 * h2));
  __align__(16) __nv_bfloat16 v[16];
  reinterpret_cast<uint4*>(v)[0] = reinterpret_cast<const uint4*>(in + i * 16)[0];
  reinterpret_cast<uint4*>(v)[1] = reinterpret_cast<const uint4*>(in + i * 16)[1];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yy = 2 * h + dy, xx = 2 * w + dx;
        if (yy < H && xx < W) y[(((size_t)b * 3 + c) * H + yy) * W + xx] = __bfloat162float(v[c * 4 + dy * 2 + dx]);
      }
}

}  // namespace

extern "C" {

int dpx_ffdnet_available(void) {
#ifdef DPX_NO_UMMA_CONV
  return 0;
#else
  return 1;
#endif
}

int dpx_ffdnet_create(int nb, int nc, dpx_ffdnet** out) {
  DPX_REQUIRE(out, "null argument");
  *out = nullptr;
#ifdef DPX_NO_UMMA_CONV
  set_error("library built without the tcgen05 convolution (CUTLASS headers not found at build time)");
  return DPX_ERR_STATE;
#else
  DPX_REQUIRE(nc == 96 && nb >= 3 && nb <= 32, "native FFDNet supports nc=96 (FFDNet-color), 3 <= nb <= 32");
  dpx_ffdnet* n = new (std::nothrow) dpx_ffdnet();
  if (!n) { set_error("out of host memory"); return DPX_ERR_NOMEM; }
  n->nb = nb; n->nc = nc;
  *out = n;
  return DPX_OK;
#endif
}

void dpx_ffdnet_destroy(dpx_ffdnet* n) {
  if (!n) return;
  for (int i = 0; i < 32; ++i) { cudaFree(n->w[i]); cudaFree(n->bias[i]); }
  cudaFree(n->act0); cudaFree(n->act1); cudaFree(n->io16); cudaFree(n->out16); cudaFree(n->ws);
  delete n;
}

// layer: 0 = head [96,13,3,3], 1..nb-2 = body [96,96,3,3], nb-1 = tail [12,96,3,3]; w, bias: device fp32 (nn.Conv2d layout)
int dpx_ffdnet_set_layer(dpx_ffdnet* n, int layer, const float* w, const float* bias, int cout, int cin, void* stream) {
  DPX_REQUIRE(n && w && bias, "null argument");
  DPX_REQUIRE(layer >= 0 && layer < n->nb, "layer %d out of range", layer);
  const bool head = layer == 0, tail = layer == n->nb - 1;
  const int cin_pad = head ? CIN_PAD : n->nc, cout_pad = tail ? COUT_TAIL_PAD : n->nc;
  DPX_REQUIRE(cin <= cin_pad && cout <= cout_pad, "layer %d: shape [%d,%d,3,3] does not fit [%d,%d]", layer, cout, cin, cout_pad, cin_pad);
  cudaStream_t s = (cudaStream_t)stream;
  const int total = cout_pad * 9 * cin_pad;
  if (!n->w[layer]) DPX_CUDA(cudaMalloc(&n->w[layer], sizeof(__nv_bfloat16) * total));
  if (!n->bias[layer]) DPX_CUDA(cudaMalloc(&n->bias[layer], sizeof(float) * cout_pad));
  k_pack_filter<<<(total + 255) / 256, 256, 0, s>>>(w, n->w[layer], cout, cin, cout_pad, cin_pad);
  DPX_LAUNCH_CHECK();
  k_pad_bias<<<1, 128, 0, s>>>(bias, n->bias[layer], cout, cout_pad);
  DPX_LAUNCH_CHECK();
  n->set[layer] = true;
  return DPX_OK;
}

// y = FFDNet(x, sigma): x, y [B,3,H,W] fp32 device; sigma device [B] (sigma_per_sample) or [1]
int dpx_ffdnet_forward(dpx_ffdnet* n, const float* x, const float* sigma, int sigma_per_sample, float* y, int B, int H, int W,
                       void* stream) {
#ifdef DPX_NO_UMMA_CONV
  set_error("library built without the tcgen05 convolution");
  return DPX_ERR_STATE;
#else
  DPX_REQUIRE(n && x && sigma && y, "null argument");
  for (int i = 0; i < n->nb; ++i) DPX_REQUIRE(n->set[i], "layer %d has no weights", i);
  cudaStream_t s = (cudaStream_t)stream;
  const int h2 = (H + 1) / 2, w2 = (W + 1) / 2;
  const size_t pix = (size_t)B * h2 * w2;
  if (n->act_cap < pix) {
    cudaFree(n->act0); cudaFree(n->act1); cudaFree(n->io16); cudaFree(n->out16);
    n->act0 = n->act1 = n->io16 = n->out16 = nullptr;
    DPX_CUDA(cudaMalloc(&n->act0, sizeof(__nv_bfloat16) * pix * n->nc));
    DPX_CUDA(cudaMalloc(&n->act1, sizeof(__nv_bfloat16) * pix * n->nc));
    DPX_CUDA(cudaMalloc(&n->io16, sizeof(__nv_bfloat16) * pix * 16));
    DPX_CUDA(cudaMalloc(&n->out16, sizeof(__nv_bfloat16) * pix * 16));
    n->act_cap = pix;
  }
  const size_t need = conv::conv_workspace(B, h2, w2);
  if (n->ws_bytes < need) {
    cudaFree(n->ws); n->ws = nullptr;
    DPX_CUDA(cudaMalloc(&n->ws, need));
    n->ws_bytes = need;
  }
  k_unshuffle_in<<<(unsigned)((pix + 255) / 256), 256, 0, s>>>(x, sigma, sigma_per_sample, n->io16, B, H, W, h2, w2);
  DPX_LAUNCH_CHECK();
  int rc = conv::conv_head(n->io16, n->w[0], n->bias[0], n->act0, B, h2, w2, n->ws, n->ws_bytes, s);
  if (rc) { set_error("tcgen05 conv (head) failed with status %d", rc); return DPX_ERR_CUDA; }
  ++g_launches;
  __nv_bfloat16 *cur = n->act0, *nxt = n->act1;
  for (int l = 1; l < n->nb - 1; ++l) {
    rc = conv::conv_body(cur, n->w[l], n->bias[l], nxt, B, h2, w2, n->ws, n->ws_bytes, s);
    if (rc) { set_error("tcgen05 conv (body %d) failed with status %d", l, rc); return DPX_ERR_CUDA; }
    ++g_launches;
    __nv_bfloat16* t = cur; cur = nxt; nxt = t;
  }
  rc = conv::conv_tail(cur, n->w[n->nb - 1], n->bias[n->nb - 1], n->out16, B, h2, w2, n->ws, n->ws_bytes, s);
  if (rc) { set_error("tcgen05 conv (tail) failed with status %d", rc); return DPX_ERR_CUDA; }
  ++g_launches;
  k_shuffle_out<<<(unsigned)((pix + 255) / 256), 256, 0, s>>>(n->out16, y, B, H, W, h2, w2);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
#endif
}

}  // extern "C"
