// dpx_ffdnet.cu — native FFDNet-color (the denoiser behind deep_prior): forward and data gradient.
//   PixelUnshuffle(2) + sigma map -> conv3x3(13->96)+ReLU -> 10 x [conv3x3(96->96)+ReLU] -> conv3x3(96->12) -> PixelShuffle(2)
//   (proxfn/pnp/denoisers/models/network_ffdnet.py:44-68)
// Activations live in HBM as channel-group-major bf16 [N][C/8][h][w][8]; the twelve convolutions run on tcgen05 tensor
// cores through the hand-written kernel of dpx_conv_tc.cuh (CTA pairs, TMEM accumulators, TMA-staged activation rows reused
// for all nine taps, filter bank resident in shared memory).  The data gradient (unrolled training with a frozen denoiser,
// BASELINE config 5) reuses the SAME kernel: d/d(input) of a 3x3 convolution is a 3x3 convolution with the transposed,
// spatially flipped filter, and the ReLU mask of the saved forward activation is applied in its epilogue.
// bf16 operands, fp32 accumulation => ~1e-2 relative accuracy: this is the FAST denoiser; the fp32 parity path keeps the
// framework convolution (dprox_b200/denoisers.py).
#include <cuda_bf16.h>

#include <new>

#include "dpx_common.cuh"
#include "dpx_conv_tc.cuh"
#include "dpx_conv_wgrad.cuh"

using namespace dpx;

namespace {

constexpr int MAXL = 32;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Tensor map of a padded channel-group-major activation tensor [N][CG][H][W+2][8] bf16.  One staged row of one channel group
// is 130 px x 16 B = 2080 contiguous bytes starting at ANY pixel; a box dimension is limited to 256 elements, so the run is
// described as {130 x 8-byte elements (65 px), 2 halves 1040 B apart} and the start pixel gets its own unit-stride (16 B)
// dimension -- the dimensions deliberately overlap in memory.  dims {130, 2, W+2, H, N*CG}, box {130, 2, 1, 1, CG}.
int make_row_map(CUtensorMap* map, const void* ptr, int N, int CG, int H, int W, int box_planes = 0) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return DPX_ERR_CUDA; }
  const cuuint64_t Wp = (cuuint64_t)convtc::row_pitch(W);
  const cuuint64_t dims[5] = {(cuuint64_t)convtc::HALO_PX, 2, Wp, (cuuint64_t)H, (cuuint64_t)N * CG};
  const cuuint64_t strides[4] = {(cuuint64_t)convtc::HALO_PX * 8, 16, Wp * 16, (cuuint64_t)H * Wp * 16};
  const cuuint32_t box[5] = {(cuuint32_t)convtc::HALO_PX, 2, 1, 1, (cuuint32_t)(box_planes ? box_planes : CG)};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d (N=%d CG=%d H=%d W=%d)", (int)r, N, CG, H, W); return DPX_ERR_CUDA; }
  return DPX_OK;
}

// elements of a padded activation buffer: rows of W + 2 pixels, plus slack for the last tile's overhang (a 130-pixel run that
// starts inside the last row may end up to 129 pixels past it)
inline size_t padded_elems(int N, int C, int H, int W) { return ((size_t)N * (C / 8) * H * convtc::row_pitch(W) + 256) * 8; }

// nn.Conv2d weight [Cout,Cin,3,3] fp32 -> the shared-memory image of the kernel: [half][tap][cg][n][8] bf16, zero padded.
// transpose = 1 builds the data-gradient filter  Wd[ci][co][ky][kx] = W[co][ci][2-ky][2-kx]  (roles of cin / cout swapped).
__global__ void k_pack_filter_tc(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cin, int cgin_pad,
                                 int cout_pad, int transpose) {
  const int nh = cout_pad / 2;
  const int total = 2 * 9 * cgin_pad * nh * 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = i % 8, n = (i / 8) % nh, cg = (i / (8 * nh)) % cgin_pad, tap = (i / (8 * nh * cgin_pad)) % 9, half = i / (8 * nh * cgin_pad * 9);
  const int ko = half * nh + n, ki = cg * 8 + c8;            // output / input channel of the (possibly transposed) filter
  const int ky = tap / 3, kx = tap % 3;
  float v = 0.f;
  if (!transpose) {
    if (ko < cout && ki < cin) v = w[(((size_t)ko * cin + ki) * 3 + ky) * 3 + kx];
  } else {
    // filter of the data gradient: its outputs are the ORIGINAL inputs (cin of them), its inputs the original outputs
    if (ko < cin && ki < cout) v = w[(((size_t)ki * cin + ko) * 3 + (2 - ky)) * 3 + (2 - kx)];
  }
  out[i] = __float2bfloat16(v);
}

// x [B,3,H,W] fp32 -> [B][2][h2][w2][8] bf16: channel c*4+dy*2+dx = x[b,c,2h+dy,2w+dx] (replicate-padded to even sizes),
// channel 12 = sigma_b, channels 13..15 = 0                                         network_ffdnet.py:54-64
__global__ void k_unshuffle_in(const float* __restrict__ x, const float* __restrict__ sigma, int sigma_per_sample,
                               __nv_bfloat16* __restrict__ out, int B, int H, int W, int h2, int w2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;          // one thread per quarter-resolution pixel
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
  const size_t plane = (size_t)h2 * convtc::row_pitch(w2);
  const size_t o = ((size_t)b * 2 * plane + (size_t)h * convtc::row_pitch(w2) + w + 1) * 8;
  *reinterpret_cast<uint4*>(out + o) = reinterpret_cast<const uint4*>(v)[0];
  *reinterpret_cast<uint4*>(out + o + plane * 8) = reinterpret_cast<const uint4*>(v)[1];
}

// [B][2][h2][w2+2][8] bf16 -> y [B,3,H,W] fp32 (PixelShuffle(2) + crop)              network_ffdnet.py:65-68
__global__ void k_shuffle_out(const __nv_bfloat16* __restrict__ in, float* __restrict__ y, int B, int H, int W, int h2, int w2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * h2 * w2;
  if (i >= total) return;
  const int w = (int)(i % w2), h = (int)((i / w2) % h2), b = (int)(i / ((size_t)w2 * h2));
  const size_t plane = (size_t)h2 * convtc::row_pitch(w2);
  const size_t o = ((size_t)b * 2 * plane + (size_t)h * convtc::row_pitch(w2) + w + 1) * 8;
  __align__(16) __nv_bfloat16 v[16];
  reinterpret_cast<uint4*>(v)[0] = *reinterpret_cast<const uint4*>(in + o);
  reinterpret_cast<uint4*>(v)[1] = *reinterpret_cast<const uint4*>(in + o + plane * 8);
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

// adjoint of k_shuffle_out: g_y [B,3,H,W] fp32 -> [B][2][h2][w2][8] bf16 (cropped positions and the 4 padding channels get 0)
__global__ void k_shuffle_out_bwd(const float* __restrict__ gy, __nv_bfloat16* __restrict__ out, int B, int H, int W, int h2, int w2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
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
        const int yy = 2 * h + dy, xx = 2 * w + dx;
        v[c * 4 + dy * 2 + dx] = __float2bfloat16((yy < H && xx < W) ? gy[(((size_t)b * 3 + c) * H + yy) * W + xx] : 0.f);
      }
  v[12] = v[13] = v[14] = v[15] = __float2bfloat16(0.f);
  const size_t plane = (size_t)h2 * convtc::row_pitch(w2);
  const size_t o = ((size_t)b * 2 * plane + (size_t)h * convtc::row_pitch(w2) + w + 1) * 8;
  *reinterpret_cast<uint4*>(out + o) = reinterpret_cast<const uint4*>(v)[0];
  *reinterpret_cast<uint4*>(out + o + plane * 8) = reinterpret_cast<const uint4*>(v)[1];
}

// adjoint of k_unshuffle_in: g16 [B][2][h2][w2][8] bf16 -> g_x [B,3,H,W] fp32 (a replicate-padded border pixel collects the
// gradients of its copies) and g_sigma[b] += sum of channel 12 (warp-shuffle + atomics)
__global__ void k_unshuffle_in_bwd(const __nv_bfloat16* __restrict__ g16, float* __restrict__ gx, float* __restrict__ gsigma,
                                   int sigma_per_sample, int B, int H, int W, int h2, int w2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * h2 * w2;
  float gs = 0.f;
  int b = 0;
  if (i < total) {
    const int w = (int)(i % w2), h = (int)((i / w2) % h2);
    b = (int)(i / ((size_t)w2 * h2));
    const size_t plane = (size_t)h2 * convtc::row_pitch(w2);
    const size_t o = ((size_t)b * 2 * plane + (size_t)h * convtc::row_pitch(w2) + w + 1) * 8;
    __align__(16) __nv_bfloat16 v[16];
    reinterpret_cast<uint4*>(v)[0] = *reinterpret_cast<const uint4*>(g16 + o);
    reinterpret_cast<uint4*>(v)[1] = *reinterpret_cast<const uint4*>(g16 + o + plane * 8);
    gs = __bfloat162float(v[12]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float g[2][2];
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) g[dy][dx] = __bfloat162float(v[c * 4 + dy * 2 + dx]);
      // clamp-duplicated positions (odd H / W, last quarter-resolution row / column) fold onto the border pixel
      const bool fold_y = 2 * h + 1 >= H, fold_x = 2 * w + 1 >= W;
      if (fold_y) { g[0][0] += g[1][0]; g[0][1] += g[1][1]; }
      if (fold_x) { g[0][0] += g[0][1]; if (!fold_y) g[1][0] += g[1][1]; }
      float* dst = gx + (((size_t)b * 3 + c) * H + 2 * h) * W + 2 * w;
      dst[0] = g[0][0];
      if (!fold_x) dst[1] = g[0][1];
      if (!fold_y) { dst[W] = g[1][0]; if (!fold_x) dst[W + 1] = g[1][1]; }
    }
  }
  if (gsigma) {
    // blocks never straddle... they may: reduce per warp only where the whole warp shares b, else fall back to per-thread atomics
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    const bool uniform = __all_sync(0xffffffffu, b == b0 || i >= total);
    if (uniform) {
      const float s = warp_sum(i < total ? gs : 0.f);
      if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(gsigma + (sigma_per_sample ? b0 : 0), s);
    } else if (i < total) {
      atomicAdd(gsigma + (sigma_per_sample ? b : 0), gs);
    }
  }
}

// ---- SPLIT (fp16 hi + 2^-11 lo') variants of the layout kernels: tensors [N][khalf][piece][kp][H][W+2][8] fp16 -------------------
__device__ __forceinline__ void split_half(float v, __half& hi, __half& lo) {
  v = fminf(fmaxf(v, -65000.f), 65000.f);
  hi = __float2half_rn(v);
  lo = __float2half_rn((v - __half2float(hi)) * 2048.f);
}
// element offset of (image n, channel group cg, piece) in a piece tensor of `cgt` channel groups with K-halves of `kp` groups
__device__ __forceinline__ size_t piece_plane(int n, int cg, int piece, int cgt, int kp) {
  const int kh = cg / kp;
  return (size_t)n * 2 * cgt + (size_t)kh * 2 * kp + (size_t)piece * kp + (cg - kh * kp);
}

// nn.Conv2d weight -> filter images of the SPLIT kernel, one per K-half: [khalf][cta half][tap][piece][kp][nh][8] fp16
__global__ void k_pack_filter_split(const float* __restrict__ w, __half* __restrict__ out, int cout, int cin, int cgin_pad, int kp,
                                    int cout_pad, int transpose) {
  const int nh = cout_pad / 2, nk = cgin_pad / kp;
  const int total = nk * 2 * 9 * 2 * kp * nh * 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int r = i;
  const int c8 = r % 8; r /= 8;
  const int n = r % nh; r /= nh;
  const int gi = r % (2 * kp); r /= 2 * kp;
  const int tap = r % 9; r /= 9;
  const int half = r % 2; r /= 2;
  const int kh = r;
  const int piece = gi / kp, cg = kh * kp + gi % kp;
  const int ko = half * nh + n, ki = cg * 8 + c8;
  const int ky = tap / 3, kx = tap % 3;
  float v = 0.f;
  if (!transpose) {
    if (ko < cout && ki < cin) v = w[(((size_t)ko * cin + ki) * 3 + ky) * 3 + kx];
  } else {
    if (ko < cin && ki < cout) v = w[(((size_t)ki * cin + ko) * 3 + (2 - ky)) * 3 + (2 - kx)];
  }
  __half hi, lo;
  split_half(v, hi, lo);
  out[i] = piece ? lo : hi;
}

// x [B,3,H,W] fp32 -> 16-channel piece tensor (cgt = 2, kp = 2): see k_unshuffle_in
__global__ void k_unshuffle_in_split(const float* __restrict__ x, const float* __restrict__ sigma, int sigma_per_sample,
                                     __half* __restrict__ out, int B, int H, int W, int h2, int w2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * h2 * w2;
  if (i >= total) return;
  const int w = (int)(i % w2), h = (int)((i / w2) % h2), b = (int)(i / ((size_t)w2 * h2));
  __align__(16) __half vh[16], vl[16];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yy = min(2 * h + dy, H - 1), xx = min(2 * w + dx, W - 1);
        split_half(x[(((size_t)b * 3 + c) * H + yy) * W + xx], vh[c * 4 + dy * 2 + dx], vl[c * 4 + dy * 2 + dx]);
      }
  split_half(sigma[sigma_per_sample ? b : 0], vh[12], vl[12]);
  vh[13] = vh[14] = vh[15] = vl[13] = vl[14] = vl[15] = __float2half_rn(0.f);
  const size_t plane = (size_t)h2 * convtc::row_pitch(w2) * 8, px = ((size_t)h * convtc::row_pitch(w2) + w + 1) * 8;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    *reinterpret_cast<uint4*>(out + piece_plane(b, g, 0, 2, 2) * plane + px) = reinterpret_cast<const uint4*>(vh)[g];
    *reinterpret_cast<uint4*>(out + piece_plane(b, g, 1, 2, 2) * plane + px) = reinterpret_cast<const uint4*>(vl)[g];
  }
}

// fp32 [B][2][h2][w2+2][8] -> y [B,3,H,W] fp32 (PixelShuffle(2) + crop)
__global__ void k_shuffle_out32(const float* __restrict__ in, float* __restrict__ y, int B, int H, int W, int h2, int w2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * h2 * w2;
  if (i >= total) return;
  const int w = (int)(i % w2), h = (int)((i / w2) % h2), b = (int)(i / ((size_t)w2 * h2));
  const size_t plane = (size_t)h2 * convtc::row_pitch(w2);
  const size_t o = ((size_t)b * 2 * plane + (size_t)h * convtc::row_pitch(w2) + w + 1) * 8;
  float v[16];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    reinterpret_cast<float4*>(v)[q] = *reinterpret_cast<const float4*>(in + o + 4 * q);
    reinterpret_cast<float4*>(v)[2 + q] = *reinterpret_cast<const float4*>(in + o + plane * 8 + 4 * q);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yy = 2 * h + dy, xx = 2 * w + dx;
        if (yy < H && xx < W) y[(((size_t)b * 3 + c) * H + yy) * W + xx] = v[c * 4 + dy * 2 + dx];
      }
}

// adjoint of k_shuffle_out32: g_y [B,3,H,W] fp32 -> 16-channel piece tensor
__global__ void k_shuffle_out_bwd_split(const float* __restrict__ gy, __half* __restrict__ out, int B, int H, int W, int h2, int w2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * h2 * w2;
  if (i >= total) return;
  const int w = (int)(i % w2), h = (int)((i / w2) % h2), b = (int)(i / ((size_t)w2 * h2));
  __align__(16) __half vh[16], vl[16];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yy = 2 * h + dy, xx = 2 * w + dx;
        split_half((yy < H && xx < W) ? gy[(((size_t)b * 3 + c) * H + yy) * W + xx] : 0.f, vh[c * 4 + dy * 2 + dx], vl[c * 4 + dy * 2 + dx]);
      }
  vh[12] = vh[13] = vh[14] = vh[15] = vl[12] = vl[13] = vl[14] = vl[15] = __float2half_rn(0.f);
  const size_t plane = (size_t)h2 * convtc::row_pitch(w2) * 8, px = ((size_t)h * convtc::row_pitch(w2) + w + 1) * 8;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    *reinterpret_cast<uint4*>(out + piece_plane(b, g, 0, 2, 2) * plane + px) = reinterpret_cast<const uint4*>(vh)[g];
    *reinterpret_cast<uint4*>(out + piece_plane(b, g, 1, 2, 2) * plane + px) = reinterpret_cast<const uint4*>(vl)[g];
  }
}

// adjoint of k_unshuffle_in for an fp32 16-channel gradient [B][2][h2][w2+2][8]: see k_unshuffle_in_bwd
__global__ void k_unshuffle_in_bwd32(const float* __restrict__ g32, float* __restrict__ gx, float* __restrict__ gsigma,
                                     int sigma_per_sample, int B, int H, int W, int h2, int w2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * h2 * w2;
  float gs = 0.f;
  int b = 0;
  if (i < total) {
    const int w = (int)(i % w2), h = (int)((i / w2) % h2);
    b = (int)(i / ((size_t)w2 * h2));
    const size_t plane = (size_t)h2 * convtc::row_pitch(w2);
    const size_t o = ((size_t)b * 2 * plane + (size_t)h * convtc::row_pitch(w2) + w + 1) * 8;
    float v[16];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      reinterpret_cast<float4*>(v)[q] = *reinterpret_cast<const float4*>(g32 + o + 4 * q);
      reinterpret_cast<float4*>(v)[2 + q] = *reinterpret_cast<const float4*>(g32 + o + plane * 8 + 4 * q);
    }
    gs = v[12];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float g[2][2];
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) g[dy][dx] = v[c * 4 + dy * 2 + dx];
      const bool fold_y = 2 * h + 1 >= H, fold_x = 2 * w + 1 >= W;
      if (fold_y) { g[0][0] += g[1][0]; g[0][1] += g[1][1]; }
      if (fold_x) { g[0][0] += g[0][1]; if (!fold_y) g[1][0] += g[1][1]; }
      float* dst = gx + (((size_t)b * 3 + c) * H + 2 * h) * W + 2 * w;
      dst[0] = g[0][0];
      if (!fold_x) dst[1] = g[0][1];
      if (!fold_y) { dst[W] = g[1][0]; if (!fold_x) dst[W + 1] = g[1][1]; }
    }
  }
  if (gsigma) {
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    const bool uniform = __all_sync(0xffffffffu, b == b0 || i >= total);
    if (uniform) {
      const float s = warp_sum(i < total ? gs : 0.f);
      if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(gsigma + (sigma_per_sample ? b0 : 0), s);
    } else if (i < total) {
      atomicAdd(gsigma + (sigma_per_sample ? b : 0), gs);
    }
  }
}

// fp32 NCHW -> piece tensor / fp32 channel-group-major -> NCHW (per-layer debug entry, SPLIT mode)
__global__ void k_nchw_to_split(const float* __restrict__ x, __half* __restrict__ out, int B, int Cc, int CG, int kp, int H, int W) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * CG * H * W * 8;
  if (i >= total) return;
  const int c8 = (int)(i % 8);
  size_t r = i / 8;
  const int w = (int)(r % W); r /= W;
  const int h = (int)(r % H); r /= H;
  const int cg = (int)(r % CG);
  const int b = (int)(r / CG);
  const int c = cg * 8 + c8;
  __half hi, lo;
  split_half(c < Cc ? x[(((size_t)b * Cc + c) * H + h) * W + w] : 0.f, hi, lo);
  const size_t plane = (size_t)H * convtc::row_pitch(W) * 8, px = ((size_t)h * convtc::row_pitch(W) + w + 1) * 8 + c8;
  out[piece_plane(b, cg, 0, CG, kp) * plane + px] = hi;
  out[piece_plane(b, cg, 1, CG, kp) * plane + px] = lo;
}
__global__ void k_c8f32_to_nchw(const float* __restrict__ in, float* __restrict__ y, int B, int Cc, int CG, int H, int W) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * Cc * H * W;
  if (i >= total) return;
  const int w = (int)(i % W);
  size_t r = i / W;
  const int h = (int)(r % H); r /= H;
  const int c = (int)(r % Cc);
  const int b = (int)(r / Cc);
  y[i] = in[((((size_t)b * CG + c / 8) * H + h) * convtc::row_pitch(W) + w + 1) * 8 + c % 8];
}

// fp32 NCHW <-> channel-group-major bf16 (per-layer debug entry)
__global__ void k_nchw_to_c8(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int Cc, int CG, int H, int W) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * CG * H * W * 8;
  if (i >= total) return;
  const int c8 = (int)(i % 8);
  size_t r = i / 8;
  const int w = (int)(r % W); r /= W;
  const int h = (int)(r % H); r /= H;
  const int cg = (int)(r % CG);
  const int b = (int)(r / CG);
  const int c = cg * 8 + c8;
  out[((((size_t)b * CG + cg) * H + h) * convtc::row_pitch(W) + w + 1) * 8 + c8] = __float2bfloat16(c < Cc ? x[(((size_t)b * Cc + c) * H + h) * W + w] : 0.f);
}
__global__ void k_c8_to_nchw(const __nv_bfloat16* __restrict__ in, float* __restrict__ y, int B, int Cc, int CG, int H, int W) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * Cc * H * W;
  if (i >= total) return;
  const int w = (int)(i % W);
  size_t r = i / W;
  const int h = (int)(r % H); r /= H;
  const int c = (int)(r % Cc);
  const int b = (int)(r / Cc);
  y[i] = __bfloat162float(in[((((size_t)b * CG + c / 8) * H + h) * convtc::row_pitch(W) + w + 1) * 8 + c % 8]);
}

template <int CGIN, int COUT>
int launch_conv(const __nv_bfloat16* in, const __nv_bfloat16* wpack, const float* bias, __nv_bfloat16* out, const __nv_bfloat16* mask,
                int relu, int N, int H, int W, cudaStream_t s) {
  using C = convtc::Cfg<CGIN, COUT>;
  static bool attr_set = false;
  if (!attr_set) {
    DPX_CUDA(cudaFuncSetAttribute(convtc::k_conv3x3_tc<CGIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    attr_set = true;
  }
  alignas(64) CUtensorMap map;
  int rc = make_row_map(&map, in, N, CGIN, H, W);
  if (rc) return rc;
  convtc::Params P;
  P.wpack = wpack; P.out = out; P.mask = mask; P.N = N; P.H = H; P.W = W; P.relu = relu;
  P.in_planes = CGIN; P.in_plane_off = 0;
  P.out16 = nullptr; P.mask16 = nullptr; P.out_kp = 1; P.pin = nullptr; P.pout = nullptr; P.out32 = nullptr;
  for (int i = 0; i < 96; ++i) P.bias[i] = (bias && i < COUT) ? bias[i] : 0.f;
  P.x_tiles = (W + convtc::TILE_PX - 1) / convtc::TILE_PX;
  P.row_blocks = (H + convtc::ROW_BLOCK - 1) / convtc::ROW_BLOCK;
  P.n_tiles = N * P.x_tiles * P.row_blocks;
  int dev = 0, sms = 0;
  DPX_CUDA(cudaGetDevice(&dev));
  DPX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int units = (P.n_tiles + 1) / 2;
  const int clusters = units < sms / 2 ? units : sms / 2;
  convtc::k_conv3x3_tc<CGIN, COUT><<<2 * clusters, C::NTHREADS, C::SMEM, s>>>(map, P);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

// One launch of the SPLIT kernel over one K-half of a layer.  `in`: piece tensor with `in_cgt` channel groups (planes per image =
// 2 in_cgt); this launch stages planes [khalf * CGIN, khalf * CGIN + CGIN) of every image (CGIN = 2 kp: hi + lo' pieces).
struct SplitOut {
  __half* out16 = nullptr; int out_kp = 1; const __half* mask16 = nullptr;
  const float* pin = nullptr; float* pout = nullptr; float* out32 = nullptr;
};
template <int CGIN, int COUT>
int launch_conv_split(const __half* in, int in_cgt, int khalf, const __half* wpack, const float* bias, const SplitOut& o, int relu,
                      int N, int H, int W, cudaStream_t s) {
  using C = convtc::Cfg<CGIN, COUT, true>;
  static bool attr_set = false;
  if (!attr_set) {
    DPX_CUDA(cudaFuncSetAttribute(convtc::k_conv3x3_tc<CGIN, COUT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    attr_set = true;
  }
  alignas(64) CUtensorMap map;
  int rc = make_row_map(&map, in, N, 2 * in_cgt, H, W, CGIN);
  if (rc) return rc;
  convtc::Params P;
  P.wpack = wpack + (size_t)khalf * 2 * (C::W_BYTES / 2);      // W_BYTES per CTA half, in bytes; fp16 elements
  P.out = nullptr; P.mask = nullptr; P.N = N; P.H = H; P.W = W; P.relu = relu;
  for (int i = 0; i < 96; ++i) P.bias[i] = (bias && i < COUT) ? bias[i] : 0.f;
  P.in_planes = 2 * in_cgt; P.in_plane_off = khalf * CGIN;
  P.out16 = o.out16; P.mask16 = o.mask16; P.out_kp = o.out_kp; P.pin = o.pin; P.pout = o.pout; P.out32 = o.out32;
  P.x_tiles = (W + convtc::TILE_PX - 1) / convtc::TILE_PX;
  P.row_blocks = (H + convtc::ROW_BLOCK - 1) / convtc::ROW_BLOCK;
  P.n_tiles = N * P.x_tiles * P.row_blocks;
  int dev = 0, sms = 0;
  DPX_CUDA(cudaGetDevice(&dev));
  DPX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int units = (P.n_tiles + 1) / 2;
  const int clusters = units < sms / 2 ? units : sms / 2;
  convtc::k_conv3x3_tc<CGIN, COUT, true><<<2 * clusters, C::NTHREADS, C::SMEM, s>>>(map, P);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

// ---- weight gradient (dpx_conv_wgrad.cuh) -------------------------------------------------------------------------------------
// dW accumulator [9][128][N] fp32 -> nn.Conv2d layout gw[co][ci][ky][kx] (+= when accumulate)
__global__ void k_wgrad_unpack(const float* __restrict__ dw, float* __restrict__ gw, int cout, int cin, int N, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * cin * 9) return;
  const int tap = i % 9, ci = (i / 9) % cin, co = i / (9 * cin);
  const float v = dw[((size_t)tap * 128 + co) * N + ci];
  gw[i] = accumulate ? gw[i] + v : v;
}
// bias gradient: db[co] = sum over pixels of gy (channel-group-major bf16 input).  grid (CG, slices): a block walks the rows of one
// channel group with a stride, keeps 8 partial sums per thread and ends with ONE atomic per channel (a block per row with a warp-level
// atomic each was 740 us per layer -- 0.8 M atomics on 96 addresses -- against 420 us for the weight gradient itself)
__global__ void k_bias_grad(const __nv_bfloat16* __restrict__ gy, float* __restrict__ db, int NB, int CG, int H, int W, int cout) {
  __shared__ float red[8][8];
  const int cg = blockIdx.x;
  const int pitch = convtc::row_pitch(W);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = blockIdx.y; r < NB * H; r += gridDim.y) {
    const int n = r / H, y = r - n * H;
    const __nv_bfloat16* src = gy + (((size_t)n * CG + cg) * H + y) * pitch * 8;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + (size_t)(x + 1) * 8);
      const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += __bfloat162float(b[e]);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float s = warp_sum(acc[e]);
    if (lane == 0) red[warp][e] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
    if (cg * 8 + threadIdx.x < cout) atomicAdd(db + cg * 8 + threadIdx.x, s);
  }
}

template <int CGG, int CGA>
int launch_wgrad(const __nv_bfloat16* gy, const __nv_bfloat16* a, float* dw, int cout, int N, int H, int W, cudaStream_t s) {
  using C = convtc::WgCfg<CGG, CGA>;
  static bool attr_set = false;
  if (!attr_set) {
    DPX_CUDA(cudaFuncSetAttribute(convtc::k_conv3x3_wgrad<CGG, CGA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    attr_set = true;
  }
  alignas(64) CUtensorMap gmap, amap;
  int rc = make_row_map(&gmap, gy, N, CGG, H, W);
  if (!rc) rc = make_row_map(&amap, a, N, CGA, H, W);
  if (rc) return rc;
  convtc::WgParams P;
  P.dw = dw; P.cout = cout; P.N = N; P.H = H; P.W = W;
  P.x_tiles = (W + convtc::TILE_PX - 1) / convtc::TILE_PX;
  P.row_blocks = (H + convtc::ROW_BLOCK - 1) / convtc::ROW_BLOCK;
  P.n_tiles = N * P.x_tiles * P.row_blocks;
  int dev = 0, sms = 0;
  DPX_CUDA(cudaGetDevice(&dev));
  DPX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ctas = P.n_tiles < sms / 2 ? P.n_tiles : sms / 2;   // x 2 tap halves = one CTA per SM
  DPX_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 9 * 128 * C::N, s));
  convtc::k_conv3x3_wgrad<CGG, CGA><<<dim3(ctas, 2), convtc::WG_THREADS, C::SMEM, s>>>(gmap, amap, P);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

}  // namespace

struct dpx_ffdnet {
  int precision = 0;                        // 0: bf16 operands (fast), 1: fp16 hi + lo' pieces (fp32-class accuracy)
  __half* ws[MAXL] = {nullptr};             // SPLIT filter images, forward / data gradient: [khalf][cta half][tap][piece][kp][nh][8]
  __half* wds[MAXL] = {nullptr};
  __half* acts[MAXL] = {nullptr};           // SPLIT activations (piece tensors)
  __half *ios = nullptr, *gas = nullptr, *gbs = nullptr;
  float *part32 = nullptr, *out32 = nullptr;
  float* dw = nullptr;                      // weight-gradient accumulator [9][128][96] fp32
  size_t acts_cap = 0;
  int n_acts = 0;
  int sbuf_B = 0, sbuf_h2 = 0, sbuf_w2 = 0;
  int nb = 0, nc = 0;
  __nv_bfloat16* w[MAXL] = {nullptr};       // forward filter, kernel image
  __nv_bfloat16* wd[MAXL] = {nullptr};      // data-gradient filter (transposed + flipped), kernel image
  float bias[MAXL][96] = {{0.f}};           // host copies: passed to the kernels as parameter constants
  bool set[MAXL] = {false};
  __nv_bfloat16 *io16 = nullptr, *out16 = nullptr;
  __nv_bfloat16* act[MAXL] = {nullptr};     // act[l] = output of layer l (l < nb - 1); two ping-pong buffers unless training
  __nv_bfloat16 *ga = nullptr, *gb = nullptr;
  size_t act_cap = 0;                       // elements per 96-channel buffer
  int n_act = 0;
  int buf_B = 0, buf_h2 = 0, buf_w2 = 0;    // shape the zero borders of the buffers were laid out for
  int saved_B = 0, saved_h2 = 0, saved_w2 = 0;
  bool saved = false;
};

namespace {

// (re)allocates the activation buffers for a [B, ., h2, w2] problem; every buffer is zero-filled whenever the shape changes,
// because the kernels rely on (and never write) the zero columns left and right of every row
int ensure_buffers(dpx_ffdnet* n, int B, int h2, int w2, bool train, cudaStream_t s) {
  const int want = train ? n->nb - 1 : 2;
  const size_t e96 = padded_elems(B, 96, h2, w2), e16 = padded_elems(B, 16, h2, w2);
  const bool shape_changed = n->buf_B != B || n->buf_h2 != h2 || n->buf_w2 != w2;
  if (n->act_cap < e96 || n->n_act < want) {
    for (int i = 0; i < MAXL; ++i) { cudaFree(n->act[i]); n->act[i] = nullptr; }
    cudaFree(n->io16); cudaFree(n->out16); cudaFree(n->ga); cudaFree(n->gb);
    n->io16 = n->out16 = n->ga = n->gb = nullptr;
    const size_t cap = e96 > n->act_cap ? e96 : n->act_cap;
    for (int i = 0; i < want; ++i) DPX_CUDA(cudaMalloc(&n->act[i], sizeof(__nv_bfloat16) * cap));
    DPX_CUDA(cudaMalloc(&n->io16, sizeof(__nv_bfloat16) * (cap / 6 + 4096)));
    DPX_CUDA(cudaMalloc(&n->out16, sizeof(__nv_bfloat16) * (cap / 6 + 4096)));
    n->act_cap = cap;
    n->n_act = want;
  } else if (!shape_changed) {
    return DPX_OK;
  }
  for (int i = 0; i < n->n_act; ++i) DPX_CUDA(cudaMemsetAsync(n->act[i], 0, sizeof(__nv_bfloat16) * e96, s));
  DPX_CUDA(cudaMemsetAsync(n->io16, 0, sizeof(__nv_bfloat16) * e16, s));
  DPX_CUDA(cudaMemsetAsync(n->out16, 0, sizeof(__nv_bfloat16) * e16, s));
  if (n->ga) {
    DPX_CUDA(cudaMemsetAsync(n->ga, 0, sizeof(__nv_bfloat16) * e96, s));
    DPX_CUDA(cudaMemsetAsync(n->gb, 0, sizeof(__nv_bfloat16) * e96, s));
  }
  n->buf_B = B; n->buf_h2 = h2; n->buf_w2 = w2;
  n->saved = false;
  return DPX_OK;
}

int run_forward(dpx_ffdnet* n, const float* x, const float* sigma, int sigma_per_sample, float* y, int B, int H, int W, bool train,
                cudaStream_t s) {
  for (int i = 0; i < n->nb; ++i) DPX_REQUIRE(n->set[i], "layer %d has no weights", i);
  const int h2 = (H + 1) / 2, w2 = (W + 1) / 2;
  const size_t pix = (size_t)B * h2 * w2;
  int rc = ensure_buffers(n, B, h2, w2, train, s);
  if (rc) return rc;
  k_unshuffle_in<<<(unsigned)((pix + 255) / 256), 256, 0, s>>>(x, sigma, sigma_per_sample, n->io16, B, H, W, h2, w2);
  DPX_LAUNCH_CHECK();
  // act[l] receives the output of layer l; without saving, two buffers alternate
  auto buf = [&](int l) { return train ? n->act[l] : n->act[l & 1]; };
  rc = launch_conv<2, 96>(n->io16, n->w[0], n->bias[0], buf(0), nullptr, 1, B, h2, w2, s);
  for (int l = 1; l < n->nb - 1 && !rc; ++l) rc = launch_conv<12, 96>(buf(l - 1), n->w[l], n->bias[l], buf(l), nullptr, 1, B, h2, w2, s);
  if (!rc) rc = launch_conv<12, 16>(buf(n->nb - 2), n->w[n->nb - 1], n->bias[n->nb - 1], n->out16, nullptr, 0, B, h2, w2, s);
  if (rc) return rc;
  k_shuffle_out<<<(unsigned)((pix + 255) / 256), 256, 0, s>>>(n->out16, y, B, H, W, h2, w2);
  DPX_LAUNCH_CHECK();
  n->saved = train;
  n->saved_B = B; n->saved_h2 = h2; n->saved_w2 = w2;
  return DPX_OK;
}


// ---- SPLIT mode: buffers, forward, data gradient ---------------------------------------------------------------------------------
int ensure_buffers_split(dpx_ffdnet* n, int B, int h2, int w2, bool train, cudaStream_t s) {
  const int want = train ? n->nb - 1 : 2;
  const size_t e96 = padded_elems(B, 2 * 96, h2, w2), e16 = padded_elems(B, 2 * 16, h2, w2);   // both pieces
  const size_t p96 = padded_elems(B, 96, h2, w2), p16 = padded_elems(B, 16, h2, w2);           // fp32 partial / result
  const bool shape_changed = n->sbuf_B != B || n->sbuf_h2 != h2 || n->sbuf_w2 != w2;
  if (n->acts_cap < e96 || n->n_acts < want) {
    for (int i = 0; i < MAXL; ++i) { cudaFree(n->acts[i]); n->acts[i] = nullptr; }
    cudaFree(n->ios); cudaFree(n->gas); cudaFree(n->gbs); cudaFree(n->part32); cudaFree(n->out32);
    n->ios = n->gas = n->gbs = nullptr; n->part32 = n->out32 = nullptr;
    const size_t cap = e96 > n->acts_cap ? e96 : n->acts_cap;
    for (int i = 0; i < want; ++i) DPX_CUDA(cudaMalloc(&n->acts[i], sizeof(__half) * cap));
    DPX_CUDA(cudaMalloc(&n->ios, sizeof(__half) * (cap / 6 + 8192)));
    DPX_CUDA(cudaMalloc(&n->part32, sizeof(float) * (cap / 2 + 4096)));
    DPX_CUDA(cudaMalloc(&n->out32, sizeof(float) * (cap / 12 + 4096)));
    n->acts_cap = cap;
    n->n_acts = want;
  } else if (!shape_changed) {
    return DPX_OK;
  }
  for (int i = 0; i < n->n_acts; ++i) DPX_CUDA(cudaMemsetAsync(n->acts[i], 0, sizeof(__half) * e96, s));
  DPX_CUDA(cudaMemsetAsync(n->ios, 0, sizeof(__half) * e16, s));
  DPX_CUDA(cudaMemsetAsync(n->part32, 0, sizeof(float) * p96, s));
  DPX_CUDA(cudaMemsetAsync(n->out32, 0, sizeof(float) * p16, s));
  if (n->gas) {
    DPX_CUDA(cudaMemsetAsync(n->gas, 0, sizeof(__half) * e96, s));
    DPX_CUDA(cudaMemsetAsync(n->gbs, 0, sizeof(__half) * e96, s));
  }
  n->sbuf_B = B; n->sbuf_h2 = h2; n->sbuf_w2 = w2;
  n->saved = false;
  return DPX_OK;
}

// a 96-channel-input layer: two launches over 48 input channels each, the fp32 partial sum handed over through part32
template <int COUT>
int conv96_split(dpx_ffdnet* n, const __half* in, const __half* w, const float* bias, SplitOut fin, int relu, int B, int h2, int w2,
                 cudaStream_t s) {
  SplitOut first;
  first.pout = n->part32;
  int rc = launch_conv_split<12, COUT>(in, 12, 0, w, nullptr, first, 0, B, h2, w2, s);
  if (rc) return rc;
  fin.pin = n->part32;
  return launch_conv_split<12, COUT>(in, 12, 1, w, bias, fin, relu, B, h2, w2, s);
}

int run_forward_split(dpx_ffdnet* n, const float* x, const float* sigma, int sigma_per_sample, float* y, int B, int H, int W, bool train,
                      cudaStream_t s) {
  for (int i = 0; i < n->nb; ++i) DPX_REQUIRE(n->set[i], "layer %d has no weights", i);
  const int h2 = (H + 1) / 2, w2 = (W + 1) / 2;
  const size_t pix = (size_t)B * h2 * w2;
  int rc = ensure_buffers_split(n, B, h2, w2, train, s);
  if (rc) return rc;
  k_unshuffle_in_split<<<(unsigned)((pix + 255) / 256), 256, 0, s>>>(x, sigma, sigma_per_sample, n->ios, B, H, W, h2, w2);
  DPX_LAUNCH_CHECK();
  auto buf = [&](int l) { return train ? n->acts[l] : n->acts[l & 1]; };
  SplitOut o;
  o.out16 = buf(0); o.out_kp = 6;
  rc = launch_conv_split<4, 96>(n->ios, 2, 0, n->ws[0], n->bias[0], o, 1, B, h2, w2, s);
  for (int l = 1; l < n->nb - 1 && !rc; ++l) {
    SplitOut f;
    f.out16 = buf(l); f.out_kp = 6;
    rc = conv96_split<96>(n, buf(l - 1), n->ws[l], n->bias[l], f, 1, B, h2, w2, s);
  }
  if (!rc) {
    SplitOut f;
    f.out32 = n->out32;
    rc = conv96_split<16>(n, buf(n->nb - 2), n->ws[n->nb - 1], n->bias[n->nb - 1], f, 0, B, h2, w2, s);
  }
  if (rc) return rc;
  k_shuffle_out32<<<(unsigned)((pix + 255) / 256), 256, 0, s>>>(n->out32, y, B, H, W, h2, w2);
  DPX_LAUNCH_CHECK();
  n->saved = train;
  n->saved_B = B; n->saved_h2 = h2; n->saved_w2 = w2;
  return DPX_OK;
}

int run_backward_split(dpx_ffdnet* n, const float* g_y, float* g_x, float* g_sigma, int sigma_per_sample, int B, int H, int W,
                       cudaStream_t s) {
  const int h2 = (H + 1) / 2, w2 = (W + 1) / 2;
  const size_t pix = (size_t)B * h2 * w2;
  if (!n->gas) {
    DPX_CUDA(cudaMalloc(&n->gas, sizeof(__half) * n->acts_cap));
    DPX_CUDA(cudaMalloc(&n->gbs, sizeof(__half) * n->acts_cap));
    DPX_CUDA(cudaMemsetAsync(n->gas, 0, sizeof(__half) * n->acts_cap, s));
    DPX_CUDA(cudaMemsetAsync(n->gbs, 0, sizeof(__half) * n->acts_cap, s));
  }
  k_shuffle_out_bwd_split<<<(unsigned)((pix + 255) / 256), 256, 0, s>>>(g_y, n->ios, B, H, W, h2, w2);
  DPX_LAUNCH_CHECK();
  __half *cur = n->gas, *nxt = n->gbs;
  SplitOut o;
  o.out16 = cur; o.out_kp = 6; o.mask16 = n->acts[n->nb - 2];
  int rc = launch_conv_split<4, 96>(n->ios, 2, 0, n->wds[n->nb - 1], nullptr, o, 0, B, h2, w2, s);
  for (int l = n->nb - 2; l >= 1 && !rc; --l) {
    SplitOut f;
    f.out16 = nxt; f.out_kp = 6; f.mask16 = n->acts[l - 1];
    rc = conv96_split<96>(n, cur, n->wds[l], nullptr, f, 0, B, h2, w2, s);
    __half* t = cur; cur = nxt; nxt = t;
  }
  if (!rc) {
    SplitOut f;
    f.out32 = n->out32;
    rc = conv96_split<16>(n, cur, n->wds[0], nullptr, f, 0, B, h2, w2, s);
  }
  if (rc) return rc;
  if (g_sigma) DPX_CUDA(cudaMemsetAsync(g_sigma, 0, sizeof(float) * (sigma_per_sample ? B : 1), s));
  k_unshuffle_in_bwd32<<<(unsigned)((pix + 255) / 256), 256, 0, s>>>(n->out32, g_x, g_sigma, sigma_per_sample, B, H, W, h2, w2);
  DPX_LAUNCH_CHECK();
  n->saved = false;
  return DPX_OK;
}

}  // namespace

extern "C" {

int dpx_ffdnet_available(void) { return 1; }

int dpx_ffdnet_create(int nb, int nc, dpx_ffdnet** out) {
  DPX_REQUIRE(out, "null argument");
  *out = nullptr;
  DPX_REQUIRE(nc == 96 && nb >= 3 && nb <= MAXL, "native FFDNet supports nc=96 (FFDNet-color), 3 <= nb <= 32");
  dpx_ffdnet* n = new (std::nothrow) dpx_ffdnet();
  if (!n) { set_error("out of host memory"); return DPX_ERR_NOMEM; }
  n->nb = nb; n->nc = nc;
  *out = n;
  return DPX_OK;
}

void dpx_ffdnet_destroy(dpx_ffdnet* n) {
  if (!n) return;
  for (int i = 0; i < MAXL; ++i) { cudaFree(n->w[i]); cudaFree(n->wd[i]); cudaFree(n->act[i]); }
  cudaFree(n->io16); cudaFree(n->out16); cudaFree(n->ga); cudaFree(n->gb);
  for (int i = 0; i < MAXL; ++i) { cudaFree(n->ws[i]); cudaFree(n->wds[i]); cudaFree(n->acts[i]); }
  cudaFree(n->ios); cudaFree(n->gas); cudaFree(n->gbs); cudaFree(n->part32); cudaFree(n->out32); cudaFree(n->dw);
  delete n;
}

// layer: 0 = head [96,13,3,3], 1..nb-2 = body [96,96,3,3], nb-1 = tail [12,96,3,3]; w, bias: device fp32 (nn.Conv2d layout)
int dpx_ffdnet_set_layer(dpx_ffdnet* n, int layer, const float* w, const float* bias, int cout, int cin, void* stream) {
  DPX_REQUIRE(n && w && bias, "null argument");
  DPX_REQUIRE(layer >= 0 && layer < n->nb, "layer %d out of range", layer);
  const bool head = layer == 0, tail = layer == n->nb - 1;
  const int cin_pad = head ? 16 : n->nc, cout_pad = tail ? 16 : n->nc;
  DPX_REQUIRE(cin <= cin_pad && cout <= cout_pad, "layer %d: shape [%d,%d,3,3] does not fit [%d,%d]", layer, cout, cin, cout_pad, cin_pad);
  cudaStream_t s = (cudaStream_t)stream;
  const int total = 9 * cin_pad * cout_pad;                    // both halves together
  if (!n->w[layer]) DPX_CUDA(cudaMalloc(&n->w[layer], sizeof(__nv_bfloat16) * total));
  if (!n->wd[layer]) DPX_CUDA(cudaMalloc(&n->wd[layer], sizeof(__nv_bfloat16) * total));
  k_pack_filter_tc<<<(total + 255) / 256, 256, 0, s>>>(w, n->w[layer], cout, cin, cin_pad / 8, cout_pad, 0);
  DPX_LAUNCH_CHECK();
  // data-gradient filter: inputs = this layer's outputs (cout_pad channels), outputs = this layer's inputs (cin_pad channels)
  k_pack_filter_tc<<<(total + 255) / 256, 256, 0, s>>>(w, n->wd[layer], cout, cin, cout_pad / 8, cin_pad, 1);
  DPX_LAUNCH_CHECK();
  {  // SPLIT images: hi + lo' pieces, one image per K-half (same element count as two bf16 images)
    const int kpf = cin_pad == 16 ? 2 : 6, kpb = cout_pad == 16 ? 2 : 6;
    if (!n->ws[layer]) DPX_CUDA(cudaMalloc(&n->ws[layer], sizeof(__half) * 2 * total));
    if (!n->wds[layer]) DPX_CUDA(cudaMalloc(&n->wds[layer], sizeof(__half) * 2 * total));
    k_pack_filter_split<<<(2 * total + 255) / 256, 256, 0, s>>>(w, n->ws[layer], cout, cin, cin_pad / 8, kpf, cout_pad, 0);
    DPX_LAUNCH_CHECK();
    k_pack_filter_split<<<(2 * total + 255) / 256, 256, 0, s>>>(w, n->wds[layer], cout, cin, cout_pad / 8, kpb, cin_pad, 1);
    DPX_LAUNCH_CHECK();
  }
  for (int i = 0; i < 96; ++i) n->bias[layer][i] = 0.f;
  DPX_CUDA(cudaMemcpyAsync(n->bias[layer], bias, sizeof(float) * cout, cudaMemcpyDeviceToHost, s));      // cold path: once per weight load
  DPX_CUDA(cudaStreamSynchronize(s));
  n->set[layer] = true;
  return DPX_OK;
}

// y = FFDNet(x, sigma): x, y [B,3,H,W] fp32 device; sigma device [B] (sigma_per_sample) or [1]
int dpx_ffdnet_set_precision(dpx_ffdnet* n, int mode) {
  DPX_REQUIRE(n, "null argument");
  DPX_REQUIRE(mode == 0 || mode == 1, "precision mode must be 0 (bf16) or 1 (fp16 hi + lo' pieces, fp32-class)");
  n->precision = mode;
  n->saved = false;
  return DPX_OK;
}

int dpx_ffdnet_forward(dpx_ffdnet* n, const float* x, const float* sigma, int sigma_per_sample, float* y, int B, int H, int W,
                       void* stream) {
  DPX_REQUIRE(n && x && sigma && y, "null argument");
  if (n->precision) return run_forward_split(n, x, sigma, sigma_per_sample, y, B, H, W, false, (cudaStream_t)stream);
  return run_forward(n, x, sigma, sigma_per_sample, y, B, H, W, false, (cudaStream_t)stream);
}

int dpx_ffdnet_forward_train(dpx_ffdnet* n, const float* x, const float* sigma, int sigma_per_sample, float* y, int B, int H, int W,
                             void* stream) {
  DPX_REQUIRE(n && x && sigma && y, "null argument");
  if (n->precision) return run_forward_split(n, x, sigma, sigma_per_sample, y, B, H, W, true, (cudaStream_t)stream);
  return run_forward(n, x, sigma, sigma_per_sample, y, B, H, W, true, (cudaStream_t)stream);
}

// shared by dpx_ffdnet_backward (gw = gb = NULL) and dpx_ffdnet_backward_params
static int backward_bf16(dpx_ffdnet* n, const float* g_y, float* g_x, float* g_sigma, int sigma_per_sample, float* const* gw,
                         float* const* gb, int B, int H, int W, cudaStream_t s) {
  const int h2 = (H + 1) / 2, w2 = (W + 1) / 2;
  const size_t pix = (size_t)B * h2 * w2;
  if (!n->ga) {
    DPX_CUDA(cudaMalloc(&n->ga, sizeof(__nv_bfloat16) * n->act_cap));
    DPX_CUDA(cudaMalloc(&n->gb, sizeof(__nv_bfloat16) * n->act_cap));
    DPX_CUDA(cudaMemsetAsync(n->ga, 0, sizeof(__nv_bfloat16) * n->act_cap, s));
    DPX_CUDA(cudaMemsetAsync(n->gb, 0, sizeof(__nv_bfloat16) * n->act_cap, s));
  }
  if (gw && !n->dw) DPX_CUDA(cudaMalloc(&n->dw, sizeof(float) * 9 * 128 * 96));
  // weight / bias gradient of layer l from its input activation and the gradient w.r.t. its pre-activation output
  auto params_of = [&](int l, const __nv_bfloat16* gy, const __nv_bfloat16* a) -> int {
    if (!gw) return DPX_OK;
    const bool head = l == 0, tail = l == n->nb - 1;
    const int cin = head ? 13 : 96, cout = tail ? 12 : 96, cin_pad = head ? 16 : 96, cout_pad = tail ? 16 : 96;
    int rc = head ? launch_wgrad<12, 2>(gy, a, n->dw, cout, B, h2, w2, s)
                  : (tail ? launch_wgrad<2, 12>(gy, a, n->dw, cout, B, h2, w2, s) : launch_wgrad<12, 12>(gy, a, n->dw, cout, B, h2, w2, s));
    if (rc) return rc;
    k_wgrad_unpack<<<(cout * cin * 9 + 255) / 256, 256, 0, s>>>(n->dw, gw[l], cout, cin, cin_pad, 0);
    DPX_LAUNCH_CHECK();
    if (gb && gb[l]) {
      DPX_CUDA(cudaMemsetAsync(gb[l], 0, sizeof(float) * cout, s));
      k_bias_grad<<<dim3(cout_pad / 8, 48), 256, 0, s>>>(gy, gb[l], B, cout_pad / 8, h2, w2, cout);
      DPX_LAUNCH_CHECK();
    }
    return DPX_OK;
  };
  // g wrt the tail's output (16 channels, 12 real)
  k_shuffle_out_bwd<<<(unsigned)((pix + 255) / 256), 256, 0, s>>>(g_y, n->out16, B, H, W, h2, w2);
  DPX_LAUNCH_CHECK();
  int rc = params_of(n->nb - 1, n->out16, n->act[n->nb - 2]);
  // tail: g wrt its input a_{nb-2}, masked by the ReLU of the layer that produced it -> g wrt that layer's pre-activation
  __nv_bfloat16 *cur = n->ga, *nxt = n->gb;
  if (!rc) rc = launch_conv<2, 96>(n->out16, n->wd[n->nb - 1], nullptr, cur, n->act[n->nb - 2], 0, B, h2, w2, s);
  for (int l = n->nb - 2; l >= 1 && !rc; --l) {
    rc = params_of(l, cur, n->act[l - 1]);
    if (!rc) rc = launch_conv<12, 96>(cur, n->wd[l], nullptr, nxt, n->act[l - 1], 0, B, h2, w2, s);
    __nv_bfloat16* t = cur; cur = nxt; nxt = t;
  }
  if (!rc) rc = params_of(0, cur, n->io16);                     // the head's input is the saved 16-channel tensor ...
  if (!rc) rc = launch_conv<12, 16>(cur, n->wd[0], nullptr, n->io16, nullptr, 0, B, h2, w2, s);     // ... overwritten here: g wrt it
  if (rc) return rc;
  if (g_sigma) DPX_CUDA(cudaMemsetAsync(g_sigma, 0, sizeof(float) * (sigma_per_sample ? B : 1), s));
  k_unshuffle_in_bwd<<<(unsigned)((pix + 255) / 256), 256, 0, s>>>(n->io16, g_x, g_sigma, sigma_per_sample, B, H, W, h2, w2);
  DPX_LAUNCH_CHECK();
  n->saved = false;                                             // io16 was reused: the saved input is gone
  return DPX_OK;
}

int dpx_ffdnet_backward(dpx_ffdnet* n, const float* g_y, float* g_x, float* g_sigma, int sigma_per_sample, int B, int H, int W,
                        void* stream) {
  DPX_REQUIRE(n && g_y && g_x, "null argument");
  const int h2 = (H + 1) / 2, w2 = (W + 1) / 2;
  DPX_REQUIRE(n->saved && n->saved_B == B && n->saved_h2 == h2 && n->saved_w2 == w2,
              "dpx_ffdnet_backward needs the activations of a matching dpx_ffdnet_forward_train call");
  cudaStream_t s = (cudaStream_t)stream;
  if (n->precision) return run_backward_split(n, g_y, g_x, g_sigma, sigma_per_sample, B, H, W, s);
  return backward_bf16(n, g_y, g_x, g_sigma, sigma_per_sample, nullptr, nullptr, B, H, W, s);
}

// Same, and the gradients w.r.t. every layer's weights and biases (training the denoiser): gw[l] device fp32 [cout,cin,3,3],
// gb[l] device fp32 [cout] (gb or its entries may be NULL).  bf16 precision only.
int dpx_ffdnet_backward_params(dpx_ffdnet* n, const float* g_y, float* g_x, float* g_sigma, int sigma_per_sample, float* const* gw,
                               float* const* gb, int B, int H, int W, void* stream) {
  DPX_REQUIRE(n && g_y && g_x && gw, "null argument");
  DPX_REQUIRE(n->precision == 0, "the weight gradient runs in the bf16 mode only");
  const int h2 = (H + 1) / 2, w2 = (W + 1) / 2;
  DPX_REQUIRE(n->saved && n->saved_B == B && n->saved_h2 == h2 && n->saved_w2 == w2,
              "dpx_ffdnet_backward_params needs the activations of a matching dpx_ffdnet_forward_train call");
  return backward_bf16(n, g_y, g_x, g_sigma, sigma_per_sample, gw, gb, B, H, W, (cudaStream_t)stream);
}

// one convolution layer on fp32 NCHW tensors (debug / per-layer parity tests): direction 0 = forward (bias, optional ReLU),
// 1 = data gradient (no bias).  x [B,cin,H,W] -> y [B,cout,H,W] with (cin, cout) the layer's logical channel counts in that
// direction.
int dpx_ffdnet_conv_layer(dpx_ffdnet* n, int layer, int direction, int relu, const float* x, float* y, int B, int H, int W, void* stream) {
  DPX_REQUIRE(n && x && y, "null argument");
  DPX_REQUIRE(layer >= 0 && layer < n->nb && n->set[layer], "layer %d out of range or without weights", layer);
  cudaStream_t s = (cudaStream_t)stream;
  const bool head = layer == 0, tail = layer == n->nb - 1;
  int cin = head ? 13 : 96, cout = tail ? 12 : 96;
  int cin_pad = head ? 16 : 96, cout_pad = tail ? 16 : 96;
  if (direction) { int t = cin; cin = cout; cout = t; t = cin_pad; cin_pad = cout_pad; cout_pad = t; }
  const size_t pix = (size_t)B * H * W;
  if (n->precision) {
    __half* a16 = nullptr;
    float *o32 = nullptr, *part = nullptr;
    const size_t ea = padded_elems(B, 2 * cin_pad, H, W), eo = padded_elems(B, cout_pad, H, W);
    DPX_CUDA(cudaMalloc(&a16, sizeof(__half) * ea));
    DPX_CUDA(cudaMalloc(&o32, sizeof(float) * eo));
    DPX_CUDA(cudaMalloc(&part, sizeof(float) * eo));
    DPX_CUDA(cudaMemsetAsync(a16, 0, sizeof(__half) * ea, s));
    DPX_CUDA(cudaMemsetAsync(o32, 0, sizeof(float) * eo, s));
    const int kp = cin_pad == 16 ? 2 : 6;
    k_nchw_to_split<<<(unsigned)((pix * cin_pad + 255) / 256), 256, 0, s>>>(x, a16, B, cin, cin_pad / 8, kp, H, W);
    DPX_LAUNCH_CHECK();
    const __half* wp = direction ? n->wds[layer] : n->ws[layer];
    const float* bp = direction ? nullptr : n->bias[layer];
    SplitOut fin;
    fin.out32 = o32;
    int rc;
    if (cin_pad == 16) rc = launch_conv_split<4, 96>(a16, 2, 0, wp, bp, fin, relu, B, H, W, s);
    else {
      SplitOut first;
      first.pout = part;
      fin.pin = part;
      if (cout_pad == 16) {
        rc = launch_conv_split<12, 16>(a16, 12, 0, wp, nullptr, first, 0, B, H, W, s);
        if (!rc) rc = launch_conv_split<12, 16>(a16, 12, 1, wp, bp, fin, relu, B, H, W, s);
      } else {
        rc = launch_conv_split<12, 96>(a16, 12, 0, wp, nullptr, first, 0, B, H, W, s);
        if (!rc) rc = launch_conv_split<12, 96>(a16, 12, 1, wp, bp, fin, relu, B, H, W, s);
      }
    }
    if (!rc) {
      k_c8f32_to_nchw<<<(unsigned)((pix * cout + 255) / 256), 256, 0, s>>>(o32, y, B, cout, cout_pad / 8, H, W);
      ++g_launches;
    }
    cudaStreamSynchronize(s);
    cudaFree(a16); cudaFree(o32); cudaFree(part);
    return rc;
  }
  __nv_bfloat16 *a = nullptr, *o = nullptr;
  DPX_CUDA(cudaMalloc(&a, sizeof(__nv_bfloat16) * padded_elems(B, cin_pad, H, W)));
  DPX_CUDA(cudaMalloc(&o, sizeof(__nv_bfloat16) * padded_elems(B, cout_pad, H, W)));
  DPX_CUDA(cudaMemsetAsync(a, 0, sizeof(__nv_bfloat16) * padded_elems(B, cin_pad, H, W), s));
  DPX_CUDA(cudaMemsetAsync(o, 0, sizeof(__nv_bfloat16) * padded_elems(B, cout_pad, H, W), s));
  k_nchw_to_c8<<<(unsigned)((pix * cin_pad + 255) / 256), 256, 0, s>>>(x, a, B, cin, cin_pad / 8, H, W);
  DPX_LAUNCH_CHECK();
  const __nv_bfloat16* wp = direction ? n->wd[layer] : n->w[layer];
  const float* bp = direction ? nullptr : n->bias[layer];
  int rc;
  if (cin_pad == 16) rc = launch_conv<2, 96>(a, wp, bp, o, nullptr, relu, B, H, W, s);
  else if (cout_pad == 16) rc = launch_conv<12, 16>(a, wp, bp, o, nullptr, relu, B, H, W, s);
  else rc = launch_conv<12, 96>(a, wp, bp, o, nullptr, relu, B, H, W, s);
  if (!rc) {
    k_c8_to_nchw<<<(unsigned)((pix * cout + 255) / 256), 256, 0, s>>>(o, y, B, cout, cout_pad / 8, H, W);
    ++g_launches;
  }
  cudaStreamSynchronize(s);
  cudaFree(a); cudaFree(o);
  return rc;
}


// Weight (and bias) gradient of one layer on fp32 NCHW tensors (per-layer parity tests): x [B,cin,H,W] = the layer's input,
// gy [B,cout,H,W] = the gradient w.r.t. its (pre-activation) output; gw [cout,cin,3,3], gb [cout] (may be NULL).  bf16 operands.
int dpx_ffdnet_wgrad_layer(dpx_ffdnet* n, int layer, const float* x, const float* gy, float* gw, float* gb, int B, int H, int W,
                           void* stream) {
  DPX_REQUIRE(n && x && gy && gw, "null argument");
  DPX_REQUIRE(layer >= 0 && layer < n->nb, "layer %d out of range", layer);
  cudaStream_t s = (cudaStream_t)stream;
  const bool head = layer == 0, tail = layer == n->nb - 1;
  const int cin = head ? 13 : 96, cout = tail ? 12 : 96, cin_pad = head ? 16 : 96, cout_pad = tail ? 16 : 96;
  const size_t pix = (size_t)B * H * W;
  __nv_bfloat16 *a = nullptr, *g = nullptr;
  float* dw = nullptr;
  DPX_CUDA(cudaMalloc(&a, sizeof(__nv_bfloat16) * padded_elems(B, cin_pad, H, W)));
  DPX_CUDA(cudaMalloc(&g, sizeof(__nv_bfloat16) * padded_elems(B, cout_pad, H, W)));
  DPX_CUDA(cudaMalloc(&dw, sizeof(float) * 9 * 128 * 96));
  DPX_CUDA(cudaMemsetAsync(a, 0, sizeof(__nv_bfloat16) * padded_elems(B, cin_pad, H, W), s));
  DPX_CUDA(cudaMemsetAsync(g, 0, sizeof(__nv_bfloat16) * padded_elems(B, cout_pad, H, W), s));
  k_nchw_to_c8<<<(unsigned)((pix * cin_pad + 255) / 256), 256, 0, s>>>(x, a, B, cin, cin_pad / 8, H, W);
  DPX_LAUNCH_CHECK();
  k_nchw_to_c8<<<(unsigned)((pix * cout_pad + 255) / 256), 256, 0, s>>>(gy, g, B, cout, cout_pad / 8, H, W);
  DPX_LAUNCH_CHECK();
  int rc;
  if (head) rc = launch_wgrad<12, 2>(g, a, dw, cout, B, H, W, s);
  else if (tail) rc = launch_wgrad<2, 12>(g, a, dw, cout, B, H, W, s);
  else rc = launch_wgrad<12, 12>(g, a, dw, cout, B, H, W, s);
  if (!rc) {
    k_wgrad_unpack<<<(cout * cin * 9 + 255) / 256, 256, 0, s>>>(dw, gw, cout, cin, cin_pad, 0);
    ++g_launches;
    if (gb) {
      DPX_CUDA(cudaMemsetAsync(gb, 0, sizeof(float) * cout, s));
      k_bias_grad<<<dim3(cout_pad / 8, 48), 256, 0, s>>>(g, gb, B, cout_pad / 8, H, W, cout);
      ++g_launches;
    }
  }
  cudaStreamSynchronize(s);
  cudaFree(a); cudaFree(g); cudaFree(dw);
  return rc;
}

}  // extern "C"
