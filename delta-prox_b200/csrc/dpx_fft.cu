// dpx_fft.cu — cuFFT engine (general sizes).
#include "dpx_fft.cuh"

#include <new>

namespace dpx {

namespace {

class CufftEngine final : public FftEngine {
 public:
  int init(const Geom& g) {
    int n[2] = {g.H, g.W};
    size_t ws1 = 0, ws2 = 0;
    DPX_CUFFT(cufftCreate(&r2c_));
    have_r2c_ = true;
    DPX_CUFFT(cufftMakePlanMany(r2c_, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, g.P, &ws1));
    DPX_CUFFT(cufftCreate(&c2r_));
    have_c2r_ = true;
    DPX_CUFFT(cufftMakePlanMany(c2r_, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, g.P, &ws2));
    ws_ = ws1 + ws2;
    return DPX_OK;
  }
  int r2c(const float* in, float2* out, cudaStream_t s) override {
    DPX_CUFFT(cufftSetStream(r2c_, s));
    DPX_CUFFT(cufftExecR2C(r2c_, const_cast<cufftReal*>(in), reinterpret_cast<cufftComplex*>(out)));
    return DPX_OK;
  }
  int c2r(float2* in, float* out, cudaStream_t s) override {
    DPX_CUFFT(cufftSetStream(c2r_, s));
    DPX_CUFFT(cufftExecC2R(c2r_, reinterpret_cast<cufftComplex*>(in), out));
    return DPX_OK;
  }
  size_t workspace_bytes() const override { return ws_; }
  void destroy() override {
    if (have_r2c_) cufftDestroy(r2c_);
    if (have_c2r_) cufftDestroy(c2r_);
    delete this;
  }

 private:
  cufftHandle r2c_{}, c2r_{};
  bool have_r2c_ = false, have_c2r_ = false;
  size_t ws_ = 0;
};

}  // namespace

int make_fused_engine(const Geom& g, FftEngine** out);   // dpx_fused_fft.cu (returns DPX_ERR_INVALID if unsupported)

int make_cufft_engine(const Geom& g, FftEngine** out) {
  *out = nullptr;
  CufftEngine* e = new (std::nothrow) CufftEngine();
  if (!e) { set_error("out of host memory"); return DPX_ERR_NOMEM; }
  int rc = e->init(g);
  if (rc) { e->destroy(); return rc; }
  *out = e;
  return DPX_OK;
}

int make_fft_engine(const Geom& g, int backend, FftEngine** out) {
  *out = nullptr;
  if (backend == 0 || backend == 2) {
    FftEngine* f = nullptr;
    int rc = make_fused_engine(g, &f);
    if (rc == DPX_OK) { *out = f; return DPX_OK; }
    if (backend == 2) return rc;
  }
  return make_cufft_engine(g, out);
}

}  // namespace dpx
