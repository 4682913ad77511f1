// dpx_fft.cuh — FFT engines behind the Fourier-diagonal x-update (proxfn/sum_square.py:150-152).
//
//   CufftEngine : batched 2-D R2C / C2R cuFFT plans, any H x W (the general path).
//   FusedEngine : hand-written sm_100a row/column FFT kernels with the spectral solve and the
//                 prox/dual update fused into them (power-of-two H, W; see dpx_fused_fft.cu).
#pragma once
#include "dpx_kernels.cuh"

namespace dpx {

class FftEngine {
 public:
  virtual ~FftEngine() {}
  // real [P,H,W] -> half spectrum [P,H,Wc] (unnormalised forward DFT, like torch.fft.fftn)
  virtual int r2c(const float* in, float2* out, cudaStream_t s) = 0;
  // half spectrum -> real (unnormalised inverse; callers fold 1/(H*W) into the spectrum). Destroys `in`.
  virtual int c2r(float2* in, float* out, cudaStream_t s) = 0;
  virtual size_t workspace_bytes() const = 0;
  virtual void destroy() = 0;
  // called whenever the plan's Fourier-diagonal constants change (standard R2C layouts): engines that keep
  // their own copies (the fused engine packs them) refresh them here
  virtual int set_constants(const float2* fb_std, const float* dq_std, int dq_batch, cudaStream_t s) {
    (void)fb_std; (void)dq_std; (void)dq_batch; (void)s;
    return DPX_OK;
  }
  virtual void reset_constants() {}
  // hint: the quadratic diagonal is the same for every channel (grey PSF) -> planes may pair across channels
  virtual void set_channel_shared(bool shared) { (void)shared; }
  // per-channel diagonal of the non-identity psi linops, standard layout [C,H,Wc] (nullptr = none)
  virtual void set_dpsi(const float* dpsi_std) { (void)dpsi_std; }
  // fully fused ADMM/HQS loop (identity psi linops, no residuals); only valid when fused() is true
  virtual bool fused() const { return false; }
  // which transform engine the last iteration call used (DPX_ENGINE_*)
  virtual int engine_mode() const { return DPX_ENGINE_CUFFT; }
  virtual int fused_iters(const Geom& g, const PsiPack& psi, bool hqs, float* x, const float2* fb, const float* dq,
                          int dq_batch, float wid, float eps, const float* rho, int rho_stride, int it0, int n_iters,
                          cudaStream_t s) {
    (void)g; (void)psi; (void)hqs; (void)x; (void)fb; (void)dq; (void)dq_batch; (void)wid; (void)eps; (void)rho;
    (void)rho_stride; (void)it0; (void)n_iters; (void)s;
    set_error("fused iterations not available on this engine");
    return DPX_ERR_STATE;
  }
  // x <- closed-form x-update of the current state (v, u) in three fused launches (staged form: an external prox follows)
  virtual int fused_xupdate(const Geom& g, const PsiPack& psi, bool hqs, float* x, float wid, float eps, const float* rho,
                            int rho_stride, int it, cudaStream_t s) {
    (void)g; (void)psi; (void)hqs; (void)x; (void)wid; (void)eps; (void)rho; (void)rho_stride; (void)it; (void)s;
    set_error("fused x-update not available on this engine");
    return DPX_ERR_STATE;
  }
};

// backend: 0 auto, 1 cuFFT, 2 fused
int make_fft_engine(const Geom& g, int backend, FftEngine** out);

}  // namespace dpx
