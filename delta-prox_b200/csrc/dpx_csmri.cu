// dpx_csmri.cu — closed-form x-update of the CS-MRI data term on complex iterates.
//
// Reference arithmetic being replaced (paths relative to /root/reference):
//   csmri._prox                       proxfn/fast/csmri.py:14-25
//       z = fft2(v); z[mask] = ((rho z + y) / (1 + rho n_psi))[mask]; return ifft2(z)
//   fft2 / ifft2 (centred, ortho)     utils/misc.py:164-193   = fftshift(fft(ifftshift(x)))
// Five launches: roll (ifftshift) -> cuFFT C2C -> masked update in un-centred coordinates (mask / y are indexed
// through the shift, 1/sqrt(HW) of both ortho transforms folded in) -> cuFFT C2C inverse -> roll (fftshift).
#include <map>
#include <mutex>
#include <tuple>

#include "dpx_common.cuh"

namespace dpx {
namespace {

constexpr int kThreads = 256;

// out[p][i][j] = scale * in[p][(i + sh) % H][(j + sw) % W]   (complex64)
__global__ void __launch_bounds__(kThreads)
    k_roll_c(const float2* __restrict__ in, float2* __restrict__ out, int H, int W, int sh, int sw, float scale) {
  const int p = blockIdx.y;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)H * W) return;
  const int i = (int)(e / W), j = (int)(e % W);
  int si = i + sh, sj = j + sw;
  if (si >= H) si -= H;
  if (sj >= W) sj -= W;
  const float2 v = in[((size_t)p * H + si) * W + sj];
  out[(size_t)p * H * W + e] = make_float2(v.x * scale, v.y * scale);
}

// Z (unnormalised spectrum, un-centred index k) <- mask[kc] ? (rho Z + sqrt(n) y[kc]) / (1 + rho n_psi) : Z,
// kc = (k + n/2) % n the centred index the reference's mask / y live in
__global__ void __launch_bounds__(kThreads)
    k_csmri_update(float2* __restrict__ Z, const float2* __restrict__ y, const float* __restrict__ mask, int mask_batch,
                   const float* __restrict__ rho, int rho_stride, float num_psi, float sqrt_n, int C, int H, int W) {
  const int p = blockIdx.y;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)H * W) return;
  const int i = (int)(e / W), j = (int)(e % W);
  int ci = i + H / 2, cj = j + W / 2;
  if (ci >= H) ci -= H;
  if (cj >= W) cj -= W;
  const size_t ce = (size_t)ci * W + cj;
  const int b = p / C;
  const float m = mask[(size_t)(mask_batch > 1 ? p : p % C) * H * W + ce];
  if (m != 0.f) {
    const float r = rho[(size_t)b * rho_stride];
    const float inv = 1.0f / (1.0f + r * num_psi);
    const float2 z = Z[(size_t)p * H * W + e], yy = y[(size_t)p * H * W + ce];
    Z[(size_t)p * H * W + e] = make_float2((r * z.x + sqrt_n * yy.x) * inv, (r * z.y + sqrt_n * yy.y) * inv);
  }
}

__global__ void __launch_bounds__(kThreads) k_real_to_complex(const float* __restrict__ x, float2* __restrict__ out, size_t n) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = make_float2(x[e], 0.f);
}
__global__ void __launch_bounds__(kThreads) k_complex_real(const float2* __restrict__ z, float* __restrict__ out, size_t n) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = z[e].x;
}

struct C2CPlan {
  cufftHandle h{};
  float2* work = nullptr;      // [P,H,W] complex scratch
};
std::mutex g_mu;
std::map<std::tuple<int, int, int, int>, C2CPlan> g_plans;     // (device, planes, H, W) -> plan (lives until process exit)

int get_plan(int P, int H, int W, C2CPlan** out) {
  int dev = 0;
  DPX_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_tuple(dev, P, H, W);
  auto it = g_plans.find(key);
  if (it == g_plans.end()) {
    C2CPlan pl;
    int n[2] = {H, W};
    size_t ws = 0;
    DPX_CUFFT(cufftCreate(&pl.h));
    DPX_CUFFT(cufftMakePlanMany(pl.h, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2C, P, &ws));
    DPX_CUDA(cudaMalloc(&pl.work, sizeof(float2) * (size_t)P * H * W));
    it = g_plans.emplace(key, pl).first;
  }
  *out = &it->second;
  return DPX_OK;
}

}  // namespace
}  // namespace dpx

using namespace dpx;

extern "C" {

int dpx_csmri_prox(const float* v, const float* y, const float* mask, int mask_batch, const float* rho, int rho_per_sample,
                   float num_psi, float* out, int batch, int channels, int height, int width, void* stream) {
  DPX_REQUIRE(v && y && mask && rho && out, "null argument");
  DPX_REQUIRE(batch > 0 && channels > 0 && height > 0 && width > 0, "bad shape");
  DPX_REQUIRE(mask_batch == 1 || mask_batch == batch, "mask_batch must be 1 or B");
  DPX_REQUIRE(v != out, "in-place call not supported");
  cudaStream_t s = (cudaStream_t)stream;
  const int P = batch * channels, H = height, W = width;
  C2CPlan* pl = nullptr;
  int rc = get_plan(P, H, W, &pl);
  if (rc) return rc;
  const dim3 grid((unsigned)(((size_t)H * W + kThreads - 1) / kThreads), (unsigned)P);
  const float2* vin = reinterpret_cast<const float2*>(v);
  float2* o = reinterpret_cast<float2*>(out);
  // ifftshift: out[i] = in[(i + n/2) % n]
  k_roll_c<<<grid, kThreads, 0, s>>>(vin, pl->work, H, W, H / 2, W / 2, 1.0f);
  DPX_LAUNCH_CHECK();
  DPX_CUFFT(cufftSetStream(pl->h, s));
  DPX_CUFFT(cufftExecC2C(pl->h, reinterpret_cast<cufftComplex*>(pl->work), reinterpret_cast<cufftComplex*>(pl->work), CUFFT_FORWARD));
  k_csmri_update<<<grid, kThreads, 0, s>>>(pl->work, reinterpret_cast<const float2*>(y), mask, mask_batch, rho,
                                           rho_per_sample ? 1 : 0, num_psi, sqrtf((float)((double)H * W)), channels, H, W);
  DPX_LAUNCH_CHECK();
  DPX_CUFFT(cufftExecC2C(pl->h, reinterpret_cast<cufftComplex*>(pl->work), reinterpret_cast<cufftComplex*>(pl->work), CUFFT_INVERSE));
  // fftshift: out[i] = in[(i - n/2) % n] = in[(i + (n - n/2)) % n]; 1/n of the two ortho transforms
  k_roll_c<<<grid, kThreads, 0, s>>>(pl->work, o, H, W, (H - H / 2) % H, (W - W / 2) % W, 1.0f / (float)((double)H * W));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int dpx_real_to_complex(const float* x, float* out, size_t n, void* stream) {
  DPX_REQUIRE(x && out, "null argument");
  k_real_to_complex<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<float2*>(out), n);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int dpx_complex_real(const float* z, float* out, size_t n, void* stream) {
  DPX_REQUIRE(z && out, "null argument");
  k_complex_real<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(z), out, n);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

}  // extern "C"
