// dpx_optics.cu — the DOE forward model that feeds the unrolled solver in every training step.
//
// Reference arithmetic being replaced (paths relative to /root/reference/dprox/contrib/optic):
//   HeightMap.get_phase_profile    doe_model.py:37-51     field = exp(i k_l dn_l h^2)
//   RGBCollimator.get_psf          doe_model.py:91-110    aperture * field -> Fresnel -> |.|^2 -> area downsample -> / sum
//   FresnelPropagator.forward      common.py:155-164      pad N/4, fft2, * H, ifft2, crop
//   area_downsampling              common.py:27-44        avg_pool2d(factor)
//   img_psf_conv (circular)        common.py:85-118       ifft2(fft2(img) * otf).real
// Forward AND backward kernels (the reference differentiates the pipeline by autograd): complex 2-D transforms are cuFFT
// C2C (sizes such as 2244 = 1496 + 2*374 are not fused-engine sizes), everything around them is fused element-wise work.
#include <map>
#include <mutex>
#include <tuple>

#include "dpx_common.cuh"

namespace dpx {
namespace {

constexpr int kThreads = 256;
inline unsigned nblk(size_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

struct C2C { cufftHandle h{}; };
std::mutex g_mu2;
std::map<std::tuple<int, int, int, int>, C2C> g_c2c;

int c2c_plan(int P, int H, int W, cufftHandle* out) {
  int dev = 0;
  DPX_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu2);
  auto key = std::make_tuple(dev, P, H, W);
  auto it = g_c2c.find(key);
  if (it == g_c2c.end()) {
    C2C pl;
    int n[2] = {H, W};
    size_t ws = 0;
    DPX_CUFFT(cufftCreate(&pl.h));
    DPX_CUFFT(cufftMakePlanMany(pl.h, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2C, P, &ws));
    it = g_c2c.emplace(key, pl).first;
  }
  *out = it->second.h;
  return DPX_OK;
}

// out[i] = scale * a[i] * (conj?) b[i % nb]
__global__ void __launch_bounds__(kThreads)
    k_cmul(const float2* __restrict__ a, const float2* __restrict__ b, float2* __restrict__ out, size_t na, size_t nb, int conj_b, float scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na) return;
  const float2 x = a[i], y = b[i % nb];
  const float yy = conj_b ? -y.y : y.y;
  out[i] = make_float2(scale * (x.x * y.x - x.y * yy), scale * (x.x * yy + x.y * y.x));
}
// out[j] = scale * sum_k g[k nb + j] * conj(a[k nb + j])     (gradient of a broadcast factor)
__global__ void __launch_bounds__(kThreads)
    k_cmul_reduce(const float2* __restrict__ g, const float2* __restrict__ a, float2* __restrict__ out, size_t nb, int batch, float scale) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nb) return;
  float re = 0.f, im = 0.f;
  for (int k = 0; k < batch; ++k) {
    const float2 x = g[(size_t)k * nb + j], y = a[(size_t)k * nb + j];
    re += x.x * y.x + x.y * y.y;
    im += x.y * y.x - x.x * y.y;
  }
  out[j] = make_float2(scale * re, scale * im);
}

// field[l][pad + y][pad + x] = aperture[y][x] * exp(i coef[l] h[y][x]^2), zero on the padded border   (M = N + 2 pad)
__global__ void __launch_bounds__(kThreads)
    k_phase_field(const float* __restrict__ h, const float* __restrict__ coef, const float* __restrict__ ap, float2* __restrict__ out, int L, int N, int pad) {
  const int M = N + 2 * pad;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)M * M) return;
  const int y = (int)(e / M) - pad, x = (int)(e % M) - pad;
  const bool in = y >= 0 && y < N && x >= 0 && x < N;
  float hv = 0.f, a = 0.f;
  if (in) { hv = h[(size_t)y * N + x]; a = ap[(size_t)y * N + x]; }
  for (int l = 0; l < L; ++l) {
    float s, c;
    sincosf(coef[l] * hv * hv, &s, &c);
    out[(size_t)l * M * M + e] = in ? make_float2(a * c, a * s) : make_float2(0.f, 0.f);
  }
}
// g_h[y][x] = sum_l 2 h coef_l * ( -g.x * f.y + g.y * f.x ),  f = aperture * exp(i phi) recomputed
__global__ void __launch_bounds__(kThreads)
    k_phase_field_bwd(const float* __restrict__ h, const float* __restrict__ coef, const float* __restrict__ ap, const float2* __restrict__ g, float* __restrict__ gh,
                      int L, int N, int pad) {
  const int M = N + 2 * pad;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)N * N) return;
  const int y = (int)(e / N), x = (int)(e % N);
  const float hv = h[e], a = ap[e];
  float acc = 0.f;
  for (int l = 0; l < L; ++l) {
    float s, c;
    sincosf(coef[l] * hv * hv, &s, &c);
    const float2 gg = g[(size_t)l * M * M + (size_t)(y + pad) * M + (x + pad)];
    acc += 2.f * hv * coef[l] * a * (-gg.x * s + gg.y * c);
  }
  gh[e] = acc;
}

// out[l][yo][xo] = scale * mean over the f x f block of |field[l][pad + f yo + dy][pad + f xo + dx]|^2
__global__ void __launch_bounds__(kThreads)
    k_abs2_pool(const float2* __restrict__ fld, float* __restrict__ out, int L, int N, int pad, int f, float scale) {
  const int M = N + 2 * pad, n = N / f;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)L * n * n) return;
  const int l = (int)(e / ((size_t)n * n)), yo = (int)((e / n) % n), xo = (int)(e % n);
  float acc = 0.f;
  for (int dy = 0; dy < f; ++dy)
    for (int dx = 0; dx < f; ++dx) {
      const float2 v = fld[(size_t)l * M * M + (size_t)(pad + f * yo + dy) * M + (pad + f * xo + dx)];
      acc += v.x * v.x + v.y * v.y;
    }
  out[e] = scale * acc / (float)(f * f);
}
// g_field = 2 * scale * g[l][y/f][x/f] / f^2 * field inside the crop, 0 on the border
__global__ void __launch_bounds__(kThreads)
    k_abs2_pool_bwd(const float2* __restrict__ fld, const float* __restrict__ g, float2* __restrict__ gf, int L, int N, int pad, int f, float scale) {
  const int M = N + 2 * pad, n = N / f;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)L * M * M) return;
  const int l = (int)(e / ((size_t)M * M)), y = (int)((e / M) % M) - pad, x = (int)(e % M) - pad;
  if (y < 0 || y >= N || x < 0 || x >= N) { gf[e] = make_float2(0.f, 0.f); return; }
  const float w = 2.f * scale * g[(size_t)l * n * n + (size_t)(y / f) * n + (x / f)] / (float)(f * f);
  const float2 v = fld[e];
  gf[e] = make_float2(w * v.x, w * v.y);
}

__global__ void __launch_bounds__(kThreads) k_sum(const float* __restrict__ x, float* __restrict__ out, size_t n) {
  __shared__ float red[32];
  float acc[1] = {0.f};
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) acc[0] += x[e];
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc[0]);
}
__global__ void __launch_bounds__(kThreads) k_dot1(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, size_t n) {
  __shared__ float red[32];
  float acc[1] = {0.f};
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) acc[0] += x[e] * y[e];
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc[0]);
}
__global__ void __launch_bounds__(kThreads) k_div_sum(const float* __restrict__ x, const float* __restrict__ s, float* __restrict__ out, size_t n) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = x[e] / s[0];
}
// g_x = (g - <g, out>) / s
__global__ void __launch_bounds__(kThreads)
    k_div_sum_bwd(const float* __restrict__ g, const float* __restrict__ s, const float* __restrict__ dotv, float* __restrict__ gx, size_t n) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) gx[e] = (g[e] - dotv[0]) / s[0];
}

}  // namespace
}  // namespace dpx

using namespace dpx;

extern "C" {

int dpx_c2c(const float* in, float* out, int planes, int height, int width, int inverse, void* stream) {
  DPX_REQUIRE(in && out && planes > 0 && height > 0 && width > 0, "bad argument");
  cufftHandle h;
  int rc = c2c_plan(planes, height, width, &h);
  if (rc) return rc;
  DPX_CUFFT(cufftSetStream(h, (cudaStream_t)stream));
  DPX_CUFFT(cufftExecC2C(h, reinterpret_cast<cufftComplex*>(const_cast<float*>(in)), reinterpret_cast<cufftComplex*>(out),
                         inverse ? CUFFT_INVERSE : CUFFT_FORWARD));
  ++g_launches;
  return DPX_OK;
}

int dpx_cmul(const float* a, const float* b, float* out, size_t n_a, size_t n_b, int conj_b, float scale, void* stream) {
  DPX_REQUIRE(a && b && out && n_b > 0 && n_a % n_b == 0, "bad argument (n_b must divide n_a)");
  k_cmul<<<nblk(n_a), kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(a), reinterpret_cast<const float2*>(b),
                                                           reinterpret_cast<float2*>(out), n_a, n_b, conj_b, scale);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int dpx_cmul_reduce(const float* g, const float* a, float* out, size_t n_b, int batch, float scale, void* stream) {
  DPX_REQUIRE(g && a && out && n_b > 0 && batch > 0, "bad argument");
  k_cmul_reduce<<<nblk(n_b), kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(g), reinterpret_cast<const float2*>(a),
                                                                  reinterpret_cast<float2*>(out), n_b, batch, scale);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int dpx_doe_field(const float* h_sqrt, const float* coef, const float* aperture, float* field, int n_lambda, int n, int pad, void* stream) {
  DPX_REQUIRE(h_sqrt && coef && aperture && field && n_lambda > 0 && n > 0 && pad >= 0, "bad argument");
  const size_t M = (size_t)n + 2 * pad;
  k_phase_field<<<nblk(M * M), kThreads, 0, (cudaStream_t)stream>>>(h_sqrt, coef, aperture, reinterpret_cast<float2*>(field), n_lambda, n, pad);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int dpx_doe_field_backward(const float* h_sqrt, const float* coef, const float* aperture, const float* g_field, float* g_h, int n_lambda,
                           int n, int pad, void* stream) {
  DPX_REQUIRE(h_sqrt && coef && aperture && g_field && g_h, "null argument");
  k_phase_field_bwd<<<nblk((size_t)n * n), kThreads, 0, (cudaStream_t)stream>>>(h_sqrt, coef, aperture, reinterpret_cast<const float2*>(g_field),
                                                                                g_h, n_lambda, n, pad);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int dpx_abs2_pool(const float* field, float* out, int n_lambda, int n, int pad, int factor, float scale, void* stream) {
  DPX_REQUIRE(field && out && factor > 0 && n % factor == 0, "bad argument (factor must divide n)");
  const size_t no = (size_t)n_lambda * (n / factor) * (n / factor);
  k_abs2_pool<<<nblk(no), kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(field), out, n_lambda, n, pad, factor, scale);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int dpx_abs2_pool_backward(const float* field, const float* g, float* g_field, int n_lambda, int n, int pad, int factor, float scale, void* stream) {
  DPX_REQUIRE(field && g && g_field && factor > 0 && n % factor == 0, "bad argument");
  const size_t M = (size_t)n + 2 * pad;
  k_abs2_pool_bwd<<<nblk((size_t)n_lambda * M * M), kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(field), g,
                                                                                        reinterpret_cast<float2*>(g_field), n_lambda, n, pad, factor, scale);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int dpx_normalize_sum(const float* x, float* out, float* sum_out, size_t n, void* stream) {
  DPX_REQUIRE(x && out && sum_out, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  DPX_CUDA(cudaMemsetAsync(sum_out, 0, sizeof(float), s));
  const unsigned g = nblk(n) < 1184u ? nblk(n) : 1184u;
  k_sum<<<g, kThreads, 0, s>>>(x, sum_out, n);
  DPX_LAUNCH_CHECK();
  k_div_sum<<<nblk(n), kThreads, 0, s>>>(x, sum_out, out, n);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int dpx_normalize_sum_backward(const float* g, const float* out, const float* sum, float* scratch1, float* g_x, size_t n, void* stream) {
  DPX_REQUIRE(g && out && sum && scratch1 && g_x, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  DPX_CUDA(cudaMemsetAsync(scratch1, 0, sizeof(float), s));
  const unsigned gb = nblk(n) < 1184u ? nblk(n) : 1184u;
  k_dot1<<<gb, kThreads, 0, s>>>(g, out, scratch1, n);
  DPX_LAUNCH_CHECK();
  k_div_sum_bwd<<<nblk(n), kThreads, 0, s>>>(g, sum, scratch1, g_x, n);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

}  // extern "C"
