// dpx_plan.cu — plan object + the extern "C" surface declared in include/dprox_b200.h.
//
// A plan is the lowered form of one compiled solver (reference: compile() -> Algorithm.__init__ ->
// least_squares.__init__, algo/primitives.py:40-67, proxfn/sum_square.py:87-110) for one shard of
// B independent problems on one GPU.  It owns the iteration-invariant constants that the reference
// recomputes every iteration (F(K^T b), sum|OTF|^2: sum_square.py:125-148) and the FFT scratch.
#include <stdarg.h>

#include <new>

#include "dpx_fft.cuh"
#include "dpx_kernels.cuh"

namespace dpx {

static thread_local char g_err[512] = "";

unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace dpx

using namespace dpx;

struct dpx_plan {
  dpx_problem_desc d;
  Geom g;
  int device = -1;
  FftEngine* fft = nullptr;
  // constants
  float2* fb = nullptr;      // F(sum_q A_q^T b_q)           [P,H,Wc]
  float* dq = nullptr;       // sum_q |OTF_q|^2  /  spatial diag
  int dq_batch = 1;
  size_t dq_bytes = 0;
  float* dpsi = nullptr;     // non-identity psi diag [C,H,Wc]
  float* ktb_sp = nullptr;   // spatial-diag numerator [P,H,W]
  float* psi_off[DPX_MAX_PSI] = {nullptr};
  float wid = 0.f;           // sum of scale^2 over identity psi terms
  bool consts_set = false;
  bool all_identity = true;
  bool has_external = false;
  // scratch
  float2* spec = nullptr;    // [P,H,Wc]
  float* t = nullptr;        // [P,H,W]
  float2* spec2 = nullptr;   // [P,H,Wc], second spectrum of the backward pass (allocated on first use)
  // host-entry state (dpx_solve_host)
  float* hx = nullptr;
  float* hv[DPX_MAX_PSI] = {nullptr};
  float* hu[DPX_MAX_PSI] = {nullptr};
  float* hsched = nullptr;
  int hsched_cap = 0;
  size_t bytes = 0;
};

namespace {

int dev_alloc(dpx_plan* p, void** ptr, size_t bytes) {
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return DPX_ERR_NOMEM;
  }
  p->bytes += bytes;
  return DPX_OK;
}

PsiPack make_pack(const dpx_plan* p, float* const* v, float* const* u, const float* const* lam, const int* lam_stride) {
  PsiPack pk;
  pk.n = p->d.n_psi;
  for (int i = 0; i < pk.n; ++i) {
    const dpx_psi_desc& s = p->d.psi[i];
    PsiTerm& t = pk.t[i];
    t.prox = s.prox; t.linop = s.linop; t.scale = s.scale; t.alpha = s.alpha; t.beta = s.beta;
    t.inv_beta = 1.0f / s.beta; t.lo = s.box_lo; t.hi = s.box_hi;
    t.v = v ? v[i] : nullptr;
    t.u = u ? u[i] : nullptr;
    t.off = p->psi_off[i];
    t.lam = lam ? lam[i] : nullptr;
    t.lam_stride = lam_stride ? lam_stride[i] : 0;
  }
  return pk;
}

bool is_admm_like(int a) { return a == DPX_ALGO_ADMM || a == DPX_ALGO_LADMM; }


int check_state_ptrs(const dpx_plan* p, const float* x, float* const* v, float* const* u) {
  DPX_REQUIRE(x != nullptr, "x is NULL");
  const int a = p->d.algo;
  if (a == DPX_ALGO_PGD) return DPX_OK;
  if (p->d.n_psi > 0) {
    DPX_REQUIRE(v != nullptr, "v is NULL");
    for (int i = 0; i < p->d.n_psi; ++i) DPX_REQUIRE(v[i] != nullptr, "v[%d] is NULL", i);
    if (a != DPX_ALGO_HQS) {
      DPX_REQUIRE(u != nullptr, "u is NULL");
      for (int i = 0; i < p->d.n_psi; ++i) DPX_REQUIRE(u[i] != nullptr, "u[%d] is NULL", i);
    }
  }
  return DPX_OK;
}

// Fused x-update for plans with stencil-gradient psi linops (TV): the vectorised stencil kernel forms
// t = sum_i s_i K_i^T (v_i - u_i), which then enters the fused engine as ONE identity term (rows: FFT of t, columns: solve with
// dq + rho (dpsi + wid), rows: inverse -> x).  (Forming the stencils inside the first row pass was measured 2.5x slower:
// 256 scalar loads per thread against 32 for a plain pass.)
int fused_xupdate_via_t(dpx_plan* p, const PsiPack& pk, bool hqs, float* x, const float* rho, int rho_stride, int it, cudaStream_t s) {
  const Geom& g = p->g;
  int rc = launch_rhs(g, pk, hqs, p->t, s);
  if (rc) return rc;
  PsiPack one;
  one.n = 1;
  memset(&one.t[0], 0, sizeof(PsiTerm));
  one.t[0].linop = DPX_LINOP_IDENTITY; one.t[0].scale = 1.f; one.t[0].alpha = one.t[0].beta = one.t[0].inv_beta = 1.f;
  one.t[0].v = p->t;
  return p->fft->fused_xupdate(g, one, /*hqs: rhs = v*/ true, x, p->wid, p->d.eps, rho, rho_stride, it, s);
}

// x <- argmin (x-update), given the rhs already in plan->t (freq) or taken from v,u (spatial)
int xupdate_freq_from_t(dpx_plan* p, float* x, RhoRef rho, cudaStream_t s) {
  const Geom& g = p->g;
  int rc = p->fft->r2c(p->t, p->spec, s);
  if (rc) return rc;
  rc = launch_spec_solve(g, p->spec, p->fb, p->dq, p->dq_batch, p->dpsi, p->wid, p->d.eps,
                         1.0f / (float)((double)g.H * g.W), rho, s);
  if (rc) return rc;
  return p->fft->c2r(p->spec, x, s);
}

}  // namespace

extern "C" {

int dpx_abi_version(void) { return DPX_ABI_VERSION; }
unsigned long long dpx_launch_count(void) { return g_launches; }
const char* dpx_last_error(void) { return g_err; }
const char* dpx_build_info(void) {
  return "libdprox_b200 abi=1 target=sm_100a cuda="
#define DPX_STR2(x) #x
#define DPX_STR(x) DPX_STR2(x)
      DPX_STR(CUDART_VERSION) " cufft=" DPX_STR(CUFFT_VERSION);
}

int dpx_plan_create(const dpx_problem_desc* d, dpx_plan** out) {
  DPX_REQUIRE(d && out, "null argument");
  *out = nullptr;
  DPX_REQUIRE(d->abi_version == DPX_ABI_VERSION, "ABI version mismatch: header %d, library %d", d->abi_version,
              DPX_ABI_VERSION);
  DPX_REQUIRE(d->batch > 0 && d->channels > 0 && d->height > 0 && d->width > 0, "bad shape [%d,%d,%d,%d]", d->batch,
              d->channels, d->height, d->width);
  DPX_REQUIRE(d->algo >= DPX_ALGO_ADMM && d->algo <= DPX_ALGO_LADMM, "unknown algo %d", d->algo);
  DPX_REQUIRE(d->xupdate == DPX_X_FREQ_DIAG || d->xupdate == DPX_X_SPATIAL_DIAG, "unknown xupdate %d", d->xupdate);
  DPX_REQUIRE(d->n_psi >= 0 && d->n_psi <= DPX_MAX_PSI, "n_psi=%d out of range", d->n_psi);
  if (d->algo == DPX_ALGO_PGD) DPX_REQUIRE(d->n_psi == 1, "PGD needs exactly one prox term (algo/pgd.py:9-26)");
  bool all_id = true, has_ext = false;
  float wid = 0.f;
  for (int i = 0; i < d->n_psi; ++i) {
    const dpx_psi_desc& s = d->psi[i];
    DPX_REQUIRE(s.prox >= DPX_PROX_NONNEG && s.prox <= DPX_PROX_ISO_TV, "psi[%d]: unknown prox %d", i, s.prox);
    DPX_REQUIRE(s.linop >= DPX_LINOP_IDENTITY && s.linop <= DPX_LINOP_GRAD_HW, "psi[%d]: unknown linop %d", i, s.linop);
    DPX_REQUIRE(s.prox != DPX_PROX_ISO_TV || s.linop == DPX_LINOP_GRAD_HW, "psi[%d]: iso-TV needs the stacked gradient", i);
    DPX_REQUIRE(!(s.linop == DPX_LINOP_GRAD_HW && s.prox == DPX_PROX_EXTERNAL), "psi[%d]: external prox on a stacked gradient", i);
    DPX_REQUIRE(s.beta != 0.f, "psi[%d]: beta must be non-zero", i);
    if (s.linop != DPX_LINOP_IDENTITY) all_id = false; else wid += s.scale * s.scale;
    if (s.prox == DPX_PROX_EXTERNAL) has_ext = true;
  }
  if (!all_id) {
    DPX_REQUIRE(d->xupdate == DPX_X_FREQ_DIAG, "grad psi linops need the Fourier-diagonal x-update");
    DPX_REQUIRE(d->algo == DPX_ALGO_ADMM || d->algo == DPX_ALGO_HQS,
                "grad psi linops are only lowered for ADMM/HQS (LADMM is inconsistent there, SURVEY App. A-6)");
  }
  dpx_plan* p = new (std::nothrow) dpx_plan();
  if (!p) { set_error("out of host memory"); return DPX_ERR_NOMEM; }
  p->d = *d;
  if (p->d.eps == 0.f) p->d.eps = 1e-7f;
  Geom& g = p->g;
  g.B = d->batch; g.C = d->channels; g.H = d->height; g.W = d->width; g.Wc = d->width / 2 + 1;
  g.P = g.B * g.C; g.plane = (size_t)g.H * g.W; g.splane = (size_t)g.H * g.Wc;
  p->all_identity = all_id; p->has_external = has_ext; p->wid = wid;
  cudaError_t ce = cudaGetDevice(&p->device);
  if (ce != cudaSuccess) { set_error("cudaGetDevice: %s", cudaGetErrorString(ce)); delete p; return DPX_ERR_CUDA; }
  int rc = DPX_OK;
  if (d->xupdate == DPX_X_FREQ_DIAG) {
    rc = dev_alloc(p, (void**)&p->spec, sizeof(float2) * g.P * g.splane);
    if (!rc) rc = make_fft_engine(g, d->fft_backend, &p->fft);
  }
  if (rc) { dpx_plan_destroy(p); return rc; }
  *out = p;
  return DPX_OK;
}

void dpx_plan_destroy(dpx_plan* p) {
  if (!p) return;
  if (p->fft) { p->fft->destroy(); }
  cudaFree(p->fb); cudaFree(p->dq); cudaFree(p->dpsi); cudaFree(p->ktb_sp); cudaFree(p->spec); cudaFree(p->spec2); cudaFree(p->t);
  cudaFree(p->hx); cudaFree(p->hsched);
  for (int i = 0; i < DPX_MAX_PSI; ++i) { cudaFree(p->psi_off[i]); cudaFree(p->hv[i]); cudaFree(p->hu[i]); }
  delete p;
}

size_t dpx_plan_workspace_bytes(const dpx_plan* p) { return p ? p->bytes + (p->fft ? p->fft->workspace_bytes() : 0) : 0; }

int dpx_plan_set_freq_constants(dpx_plan* p, const float* ktb, const float* dq, int dq_batch, const float* dpsi,
                                void* stream) {
  DPX_REQUIRE(p, "null plan");
  DPX_REQUIRE(p->d.xupdate == DPX_X_FREQ_DIAG, "plan is not FREQ_DIAG");
  DPX_REQUIRE(dq_batch == 1 || dq_batch == p->g.B, "dq_batch must be 1 or B");
  cudaStream_t s = (cudaStream_t)stream;
  const Geom& g = p->g;
  if (!p->fb) { int rc = dev_alloc(p, (void**)&p->fb, sizeof(float2) * g.P * g.splane); if (rc) return rc; }
  if (!p->t) { int rc = dev_alloc(p, (void**)&p->t, sizeof(float) * g.P * g.plane); if (rc) return rc; }
  if (ktb) {
    int rc = p->fft->r2c(ktb, p->fb, s);
    if (rc) return rc;
  } else {
    DPX_CUDA(cudaMemsetAsync(p->fb, 0, sizeof(float2) * g.P * g.splane, s));
  }
  // (re)allocate only when the size changes: cudaFree synchronises the device, which would serialise callers that
  // refresh the constants of several plans on different streams
  if (dq) {
    const size_t n = sizeof(float) * (size_t)dq_batch * g.C * g.splane;
    if (!p->dq || p->dq_bytes != n) {
      cudaFree(p->dq); p->dq = nullptr;
      int rc = dev_alloc(p, (void**)&p->dq, n);
      if (rc) return rc;
      p->dq_bytes = n;
    }
    DPX_CUDA(cudaMemcpyAsync(p->dq, dq, n, cudaMemcpyDeviceToDevice, s));
    p->dq_batch = dq_batch;
  } else if (p->dq) {
    cudaFree(p->dq); p->dq = nullptr; p->dq_bytes = 0;
  }
  if (dpsi) {
    const size_t n = sizeof(float) * (size_t)g.C * g.splane;
    if (!p->dpsi) {
      int rc = dev_alloc(p, (void**)&p->dpsi, n);
      if (rc) return rc;
    }
    DPX_CUDA(cudaMemcpyAsync(p->dpsi, dpsi, n, cudaMemcpyDeviceToDevice, s));
  } else if (p->dpsi) {
    cudaFree(p->dpsi); p->dpsi = nullptr;
  }
  {
    if (!p->dq) p->fft->reset_constants();
    p->fft->set_dpsi(p->dpsi);
    int rc = p->fft->set_constants(p->fb, p->dq, p->dq ? p->dq_batch : 1, s);
    if (rc) return rc;
  }
  p->consts_set = true;
  return DPX_OK;
}

int dpx_plan_set_rhs(dpx_plan* p, const float* ktb, void* stream) {
  DPX_REQUIRE(p, "null plan");
  DPX_REQUIRE(p->consts_set, "constants not set yet (call dpx_plan_set_*_constants first)");
  cudaStream_t s = (cudaStream_t)stream;
  const Geom& g = p->g;
  if (p->d.xupdate == DPX_X_SPATIAL_DIAG) {
    const size_t n = sizeof(float) * g.P * g.plane;
    if (ktb) {
      if (!p->ktb_sp) { int rc = dev_alloc(p, (void**)&p->ktb_sp, n); if (rc) return rc; }
      DPX_CUDA(cudaMemcpyAsync(p->ktb_sp, ktb, n, cudaMemcpyDeviceToDevice, s));
    } else if (p->ktb_sp) {
      DPX_CUDA(cudaMemsetAsync(p->ktb_sp, 0, n, s));
    }
    return DPX_OK;
  }
  if (ktb) {
    int rc = p->fft->r2c(ktb, p->fb, s);
    if (rc) return rc;
  } else {
    DPX_CUDA(cudaMemsetAsync(p->fb, 0, sizeof(float2) * g.P * g.splane, s));
  }
  return p->fft->set_constants(p->fb, nullptr, p->dq_batch, s);      // engines re-pack F(K^T b) only
}

int dpx_plan_set_hint(dpx_plan* p, int hint, int value) {
  DPX_REQUIRE(p, "null plan");
  DPX_REQUIRE(hint == DPX_HINT_CHANNEL_SHARED_DIAG, "unknown hint %d", hint);
  if (p->fft) p->fft->set_channel_shared(value != 0);
  return DPX_OK;
}

int dpx_plan_set_rhs_spectral(dpx_plan* p, const float* b, const float* otf, int otf_batch, float scale, void* stream) {
  DPX_REQUIRE(p && b && otf, "null argument");
  DPX_REQUIRE(p->consts_set, "constants not set yet (call dpx_plan_set_*_constants first)");
  DPX_REQUIRE(p->d.xupdate == DPX_X_FREQ_DIAG && p->fft && p->fb, "dpx_plan_set_rhs_spectral needs a FREQ_DIAG plan");
  DPX_REQUIRE(otf_batch == 1 || otf_batch == p->g.B, "otf_batch must be 1 or B");
  cudaStream_t s = (cudaStream_t)stream;
  int rc = p->fft->r2c(b, p->fb, s);                       // F(b) ...
  if (!rc) rc = launch_mul_otf(p->g, p->fb, (const float2*)otf, otf_batch, /*conj=*/true, scale, s);   // ... times scale * conj(OTF)
  if (!rc) rc = p->fft->set_constants(p->fb, nullptr, p->dq_batch, s);
  return rc;
}

int dpx_plan_set_spatial_constants(dpx_plan* p, const float* ktb, const float* dq, int dq_batch, void* stream) {
  DPX_REQUIRE(p, "null plan");
  DPX_REQUIRE(p->d.xupdate == DPX_X_SPATIAL_DIAG, "plan is not SPATIAL_DIAG");
  DPX_REQUIRE(dq_batch == 1 || dq_batch == p->g.B, "dq_batch must be 1 or B");
  cudaStream_t s = (cudaStream_t)stream;
  const Geom& g = p->g;
  if (!p->t) { int rc = dev_alloc(p, (void**)&p->t, sizeof(float) * g.P * g.plane); if (rc) return rc; }
  cudaFree(p->dq); p->dq = nullptr;
  cudaFree(p->ktb_sp); p->ktb_sp = nullptr;
  if (ktb) {
    const size_t n = sizeof(float) * g.P * g.plane;
    int rc = dev_alloc(p, (void**)&p->ktb_sp, n);
    if (rc) return rc;
    DPX_CUDA(cudaMemcpyAsync(p->ktb_sp, ktb, n, cudaMemcpyDeviceToDevice, s));
  }
  if (dq) {
    const size_t n = sizeof(float) * (size_t)dq_batch * g.C * g.plane;
    int rc = dev_alloc(p, (void**)&p->dq, n);
    if (rc) return rc;
    DPX_CUDA(cudaMemcpyAsync(p->dq, dq, n, cudaMemcpyDeviceToDevice, s));
    p->dq_batch = dq_batch;
  }
  p->consts_set = true;
  return DPX_OK;
}

int dpx_plan_set_spatial_psi_diag(dpx_plan* p, const float* dpsi, void* stream) {
  DPX_REQUIRE(p, "null plan");
  DPX_REQUIRE(p->d.xupdate == DPX_X_SPATIAL_DIAG, "plan is not SPATIAL_DIAG");
  if (!dpsi) { cudaFree(p->dpsi); p->dpsi = nullptr; return DPX_OK; }
  const size_t n = sizeof(float) * (size_t)p->g.C * p->g.plane;
  if (!p->dpsi) {
    int rc = dev_alloc(p, (void**)&p->dpsi, n);
    if (rc) return rc;
  }
  DPX_CUDA(cudaMemcpyAsync(p->dpsi, dpsi, n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return DPX_OK;
}

int dpx_plan_engine_mode(const dpx_plan* p) {
  if (!p || !p->fft) return DPX_ENGINE_NONE;
  return p->fft->engine_mode();
}

int dpx_plan_set_psi_offset(dpx_plan* p, int i, const float* c, void* stream) {
  DPX_REQUIRE(p, "null plan");
  DPX_REQUIRE(i >= 0 && i < p->d.n_psi, "psi index %d out of range", i);
  DPX_REQUIRE(!c || p->d.psi[i].linop != DPX_LINOP_GRAD_HW, "psi[%d]: offsets are not supported on a stacked gradient", i);
  cudaStream_t s = (cudaStream_t)stream;
  if (!c) { cudaFree(p->psi_off[i]); p->psi_off[i] = nullptr; return DPX_OK; }
  const size_t n = sizeof(float) * p->g.P * p->g.plane;
  if (!p->psi_off[i]) {
    int rc = dev_alloc(p, (void**)&p->psi_off[i], n);
    if (rc) return rc;
  }
  DPX_CUDA(cudaMemcpyAsync(p->psi_off[i], c, n, cudaMemcpyDeviceToDevice, s));
  return DPX_OK;
}

int dpx_init_state(dpx_plan* p, const float* x, float* const* v, float* const* u, void* stream) {
  DPX_REQUIRE(p, "null plan");
  if (p->d.algo == DPX_ALGO_PGD || p->d.n_psi == 0) return DPX_OK;
  int rc = check_state_ptrs(p, x, v, u);
  if (rc) return rc;
  const bool with_u = p->d.algo != DPX_ALGO_HQS;
  return launch_init_state(p->g, make_pack(p, v, u, nullptr, nullptr), x, with_u, (cudaStream_t)stream);
}

int dpx_stage_xupdate(dpx_plan* p, float* x, float* const* v, float* const* u, const float* rho, int rho_stride, int it,
                      void* stream) {
  DPX_REQUIRE(p, "null plan");
  DPX_REQUIRE(p->consts_set, "constants not set (dpx_plan_set_*_constants)");
  DPX_REQUIRE(rho, "rho schedule is NULL");
  int rc = check_state_ptrs(p, x, v, u);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const Geom& g = p->g;
  const RhoRef rr{rho, rho_stride, it};
  const int a = p->d.algo;
  const bool hqs = a == DPX_ALGO_HQS;
  if (a == DPX_ALGO_PGD) {
    float* dst = (v && v[0]) ? v[0] : p->t;
    if (p->d.xupdate == DPX_X_FREQ_DIAG) {
      rc = p->fft->r2c(x, p->spec, s);
      if (!rc) rc = launch_spec_pgd(g, p->spec, p->fb, p->dq, p->dq_batch, 1.0f / (float)((double)g.H * g.W), rr, s);
      if (!rc) rc = p->fft->c2r(p->spec, dst, s);
      return rc;
    }
    DPX_REQUIRE(p->ktb_sp && p->dq, "spatial PGD needs ktb and dq");
    return launch_pgd_spatial_step(g, x, p->ktb_sp, p->dq, p->dq_batch, rr, dst, s);
  }
  DPX_REQUIRE(a != DPX_ALGO_ADMM_VXU, "ADMM_vxu is only lowered through dpx_iters (or compose it with dpx_xsolve)");
  PsiPack pk = make_pack(p, v, u, nullptr, nullptr);
  if (p->d.xupdate == DPX_X_SPATIAL_DIAG)
    return launch_spatial_xupdate(g, pk, hqs, false, p->ktb_sp, p->dq, p->dq_batch, p->dpsi, p->wid, p->d.eps, p->d.eps_delta != 0, rr, x, s);
  // fused engine: rhs + row FFT, column FFT + solve + inverse, inverse row FFT -> x  (3 launches, 34 B/element
  // instead of the 6 launches / ~60 B/element of rhs kernel + cuFFT R2C + solve + cuFFT C2R)
  if (p->fft->fused() && pk.n > 0 && (a == DPX_ALGO_ADMM || a == DPX_ALGO_LADMM || hqs)) {
    if (p->all_identity) return p->fft->fused_xupdate(g, pk, hqs, x, p->wid, p->d.eps, rho, rho_stride, it, s);
    return fused_xupdate_via_t(p, pk, hqs, x, rho, rho_stride, it, s);
  }
  if (pk.n > 0) {
    rc = launch_rhs(g, pk, hqs, p->t, s);
    if (rc) return rc;
  } else {
    DPX_CUDA(cudaMemsetAsync(p->t, 0, sizeof(float) * g.P * g.plane, s));
  }
  return xupdate_freq_from_t(p, x, rr, s);
}

int dpx_xsolve(dpx_plan* p, const float* t, const float* rho, int rho_stride, int it, float* x, void* stream) {
  DPX_REQUIRE(p && t && rho && x, "null argument");
  DPX_REQUIRE(p->consts_set, "constants not set (dpx_plan_set_*_constants)");
  cudaStream_t s = (cudaStream_t)stream;
  const Geom& g = p->g;
  const RhoRef rr{rho, rho_stride, it};
  if (p->d.xupdate == DPX_X_SPATIAL_DIAG) {
    PsiPack one;
    one.n = 1;
    memset(&one.t[0], 0, sizeof(PsiTerm));
    one.t[0].scale = 1.f;
    one.t[0].v = const_cast<float*>(t);
    return launch_spatial_xupdate(g, one, /*hqs=*/true, false, p->ktb_sp, p->dq, p->dq_batch, p->dpsi, p->wid, p->d.eps, p->d.eps_delta != 0, rr, x, s);
  }
  if (p->fft->fused() && aligned16(t) && aligned16(x)) {      // rows: FFT of t; columns: solve; rows: inverse -> x (3 fused launches)
    PsiPack one;
    one.n = 1;
    memset(&one.t[0], 0, sizeof(PsiTerm));
    one.t[0].linop = DPX_LINOP_IDENTITY; one.t[0].scale = 1.f; one.t[0].alpha = one.t[0].beta = one.t[0].inv_beta = 1.f;
    one.t[0].v = const_cast<float*>(t);
    return p->fft->fused_xupdate(g, one, /*hqs: rhs = v*/ true, x, p->wid, p->d.eps, rho, rho_stride, it, s);
  }
  int rc = p->fft->r2c(t, p->spec, s);
  if (!rc) rc = launch_spec_solve(g, p->spec, p->fb, p->dq, p->dq_batch, p->dpsi, p->wid, p->d.eps,
                                  1.0f / (float)((double)g.H * g.W), rr, s);
  if (!rc) rc = p->fft->c2r(p->spec, x, s);
  return rc;
}

int dpx_xsolve_backward(dpx_plan* p, const float* g, const float* x, const float* rho, int rho_stride, int it, float* g_ktb,
                        float* g_rho, void* stream) {
  DPX_REQUIRE(p && g && x && rho && g_ktb, "null argument");
  DPX_REQUIRE(p->consts_set, "constants not set (dpx_plan_set_*_constants)");
  cudaStream_t s = (cudaStream_t)stream;
  const Geom& gm = p->g;
  if (p->d.xupdate == DPX_X_SPATIAL_DIAG)
    return launch_spatial_xupdate_bwd(gm, g, x, p->ktb_sp, p->dq, p->dq_batch, p->dpsi, p->wid, p->d.eps, p->d.eps_delta != 0,
                                      RhoRef{rho, rho_stride, it}, g_ktb, g_rho, rho_stride ? 1 : 0, s);
  DPX_REQUIRE(p->fft, "plan has no FFT engine");
  if (!p->spec2) {
    int rc = dev_alloc(p, (void**)&p->spec2, sizeof(float2) * gm.P * gm.splane);
    if (rc) return rc;
  }
  int rc = p->fft->r2c(g, p->spec, s);
  if (!rc && g_rho) rc = p->fft->r2c(x, p->spec2, s);
  if (!rc) rc = launch_spec_solve_bwd(gm, p->spec, p->spec2, p->fb, p->dq, p->dq_batch, p->dpsi, p->wid, p->d.eps,
                                      1.0f / (float)((double)gm.H * gm.W), RhoRef{rho, rho_stride, it}, g_rho, rho_stride ? 1 : 0, s);
  if (!rc) rc = p->fft->c2r(p->spec, g_ktb, s);
  return rc;
}

int dpx_prox_backward(int prox_kind, const float* v, const float* lam, int lam_per_sample, float alpha, float beta,
                      float box_lo, float box_hi, const float* offset, const float* g, float* g_v, float* g_lam, int batch,
                      size_t per_sample, void* stream) {
  DPX_REQUIRE(v && lam && g && g_v, "null argument");
  DPX_REQUIRE(prox_kind >= DPX_PROX_NONNEG && prox_kind <= DPX_PROX_BOX, "prox kind %d has no native backward", prox_kind);
  DPX_REQUIRE(beta != 0.f, "beta must be non-zero");
  const ProxSpec ps{prox_kind, alpha, beta, 1.0f / beta, box_lo, box_hi};
  return launch_prox_bwd(ps, v, lam, lam_per_sample ? 1 : 0, offset, g, g_v, g_lam, batch, per_sample, (cudaStream_t)stream);
}

int dpx_stage_prox(dpx_plan* p, float* x, float* const* v, float* const* u, const float* const* lam,
                   const int* lam_stride, int it, void* stream) {
  DPX_REQUIRE(p, "null plan");
  int rc = check_state_ptrs(p, x, v, u);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int a = p->d.algo;
  if (p->d.n_psi == 0) return DPX_OK;
  DPX_REQUIRE(lam && lam_stride, "lam schedule is NULL");
  if (a == DPX_ALGO_PGD) {
    const dpx_psi_desc& sd = p->d.psi[0];
    DPX_REQUIRE(sd.prox != DPX_PROX_EXTERNAL, "external prox: evaluate it in the caller");
    const float* src = (v && v[0]) ? v[0] : p->t;
    const ProxSpec ps{sd.prox, sd.alpha, sd.beta, 1.0f / sd.beta, sd.box_lo, sd.box_hi};
    return launch_prox_apply(ps, src, lam[0], lam_stride[0], it, p->psi_off[0], x, p->g.B, (size_t)p->g.C * p->g.plane, s);
  }
  PsiPack pk = make_pack(p, v, u, lam, lam_stride);
  DPX_REQUIRE(a != DPX_ALGO_ADMM_VXU, "ADMM_vxu is only lowered through dpx_iters");
  return launch_prox_dual(p->g, pk, x, a == DPX_ALGO_HQS, /*skip_external=*/true, it, nullptr, RhoRef{nullptr, 0, 0},
                          nullptr, s);
}

int dpx_stage_dual_external(dpx_plan* p, int i, const float* w, const float* v_new, float* v_i, float* u_i, void* stream) {
  DPX_REQUIRE(p && w && v_new && v_i, "null argument");
  DPX_REQUIRE(i >= 0 && i < p->d.n_psi, "psi index out of range");
  return launch_dual_external(w, v_new, v_i, p->d.algo == DPX_ALGO_HQS ? nullptr : u_i, p->g.P * p->g.plane,
                              (cudaStream_t)stream);
}

int dpx_iters(dpx_plan* p, float* x, float* const* v, float* const* u, const float* rho, int rho_stride,
              const float* const* lam, const int* lam_stride, int it0, int n_iters, float* resid, void* stream) {
  DPX_REQUIRE(p, "null plan");
  DPX_REQUIRE(p->consts_set, "constants not set (dpx_plan_set_*_constants)");
  DPX_REQUIRE(!p->has_external, "plan has an external prox term: drive it with dpx_stage_*");
  DPX_REQUIRE(rho, "rho schedule is NULL");
  DPX_REQUIRE(n_iters >= 0 && it0 >= 0, "bad iteration range");
  int rc = check_state_ptrs(p, x, v, u);
  if (rc) return rc;
  if (p->d.n_psi > 0) DPX_REQUIRE(lam && lam_stride, "lam schedule is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  const Geom& g = p->g;
  const int a = p->d.algo;
  const bool hqs = a == DPX_ALGO_HQS;
  const bool freq = p->d.xupdate == DPX_X_FREQ_DIAG;
  const float inv_n = 1.0f / (float)((double)g.H * g.W);
  PsiPack pk = make_pack(p, v, u, lam, lam_stride);
  if (resid) DPX_CUDA(cudaMemsetAsync(resid, 0, sizeof(float) * (size_t)n_iters * g.B * 4, s));

  bool vec_ok = aligned16(x);
  for (int i = 0; i < pk.n; ++i)
    vec_ok = vec_ok && aligned16(pk.t[i].v) && (hqs || aligned16(pk.t[i].u)) && (!pk.t[i].off || aligned16(pk.t[i].off));
  if (freq && p->fft->fused() && (is_admm_like(a) || hqs) && p->all_identity && !resid && vec_ok && pk.n > 0) {
    return p->fft->fused_iters(g, pk, hqs, x, p->fb, p->dq, p->dq_batch, p->wid, p->d.eps, rho, rho_stride, it0, n_iters, s);
  }

  if (freq && p->fft->fused() && (is_admm_like(a) || hqs) && !p->all_identity && pk.n > 0 && !resid && vec_ok) {
    // TV-type terms (stencil gradients): stencil rhs kernel, fused x-update (rows: FFT of t, columns, rows: inverse -> x),
    // stencil prox/dual kernel: 5 launches, the three transforms without cuFFT's extra passes over the data
    for (int k = 0; k < n_iters; ++k) {
      rc = fused_xupdate_via_t(p, pk, hqs, x, rho, rho_stride, it0 + k, s);
      if (!rc) rc = launch_prox_dual(g, pk, x, hqs, false, it0 + k, nullptr, RhoRef{rho, rho_stride, it0 + k}, nullptr, s);
      if (rc) return rc;
    }
    return DPX_OK;
  }
  bool t_valid = false;
  for (int k = 0; k < n_iters; ++k) {
    const int it = it0 + k;
    const RhoRef rr{rho, rho_stride, it};
    float* res_it = resid ? resid + (size_t)k * g.B * 4 : nullptr;
    if (a == DPX_ALGO_PGD) {
      if (freq) {
        rc = p->fft->r2c(x, p->spec, s);
        if (!rc) rc = launch_spec_pgd(g, p->spec, p->fb, p->dq, p->dq_batch, inv_n, rr, s);
        if (!rc) rc = p->fft->c2r(p->spec, p->t, s);
      } else {
        DPX_REQUIRE(p->ktb_sp && p->dq, "spatial PGD needs ktb and dq");
        rc = launch_pgd_spatial_step(g, x, p->ktb_sp, p->dq, p->dq_batch, rr, p->t, s);
      }
      if (rc) return rc;
      {  // x = prox(t, lam)   (pgd.py:42)
        const dpx_psi_desc& sd = p->d.psi[0];
        const ProxSpec ps{sd.prox, sd.alpha, sd.beta, 1.0f / sd.beta, sd.box_lo, sd.box_hi};
        rc = launch_prox_apply(ps, p->t, lam[0], lam_stride[0], it, p->psi_off[0], x, g.B, (size_t)g.C * g.plane, s);
      }
      if (rc) return rc;
      continue;
    }
    if (a == DPX_ALGO_ADMM_VXU) {
      rc = launch_vxu_prox(g, pk, x, it, freq ? p->t : nullptr, s);
      if (rc) return rc;
      if (freq) rc = xupdate_freq_from_t(p, x, rr, s);
      else rc = launch_spatial_xupdate(g, pk, false, true, p->ktb_sp, p->dq, p->dq_batch, p->dpsi, p->wid, p->d.eps, p->d.eps_delta != 0, rr, x, s);
      if (!rc) rc = launch_vxu_dual(g, pk, x, s);
      if (rc) return rc;
      continue;
    }
    // ADMM / LADMM(identity) / HQS
    if (freq) {
      if (!t_valid) {
        if (pk.n > 0) rc = launch_rhs(g, pk, hqs, p->t, s);
        else { DPX_CUDA(cudaMemsetAsync(p->t, 0, sizeof(float) * g.P * g.plane, s)); }
        if (rc) return rc;
      }
      rc = xupdate_freq_from_t(p, x, rr, s);
    } else {
      rc = launch_spatial_xupdate(g, pk, hqs, false, p->ktb_sp, p->dq, p->dq_batch, p->dpsi, p->wid, p->d.eps, p->d.eps_delta != 0, rr, x, s);
    }
    if (rc) return rc;
    if (pk.n > 0) {
      const bool fuse = freq && p->all_identity && (k + 1 < n_iters);
      rc = launch_prox_dual(g, pk, x, hqs, false, it, fuse ? p->t : nullptr, rr, res_it, s);
      if (rc) return rc;
      t_valid = fuse;
    }
  }
  return DPX_OK;
}

int dpx_spectral_filter(dpx_plan* p, const float* x, const float* otf, int otf_batch, int conjugate, float* y,
                        void* stream) {
  DPX_REQUIRE(p && x && otf && y, "null argument");
  DPX_REQUIRE(p->fft, "plan has no FFT engine (SPATIAL_DIAG plan)");
  DPX_REQUIRE(otf_batch == 1 || otf_batch == p->g.B, "otf_batch must be 1 or B");
  cudaStream_t s = (cudaStream_t)stream;
  const Geom& g = p->g;
  int rc = p->fft->r2c(x, p->spec, s);
  if (!rc) rc = launch_mul_otf(g, p->spec, (const float2*)otf, otf_batch, conjugate != 0, 1.0f / (float)((double)g.H * g.W), s);
  if (!rc) rc = p->fft->c2r(p->spec, y, s);
  return rc;
}

int dpx_prox_apply(int prox_kind, const float* v, const float* lam, int lam_per_sample, float alpha, float beta,
                   float box_lo, float box_hi, const float* offset, float* out, int batch, size_t per_sample,
                   void* stream) {
  DPX_REQUIRE(v && lam && out, "null argument");
  DPX_REQUIRE((prox_kind >= DPX_PROX_NONNEG && prox_kind <= DPX_PROX_BOX) || prox_kind == DPX_PROX_ISO_TV,
              "prox kind %d has no native kernel", prox_kind);
  DPX_REQUIRE(beta != 0.f, "beta must be non-zero");
  const ProxSpec ps{prox_kind, alpha, beta, 1.0f / beta, box_lo, box_hi};
  return launch_prox_apply(ps, v, lam, lam_per_sample ? 1 : 0, 0, offset, out, batch, per_sample, (cudaStream_t)stream);
}

int dpx_lincomb(float* out, const float* a, const float* x, const float* b, const float* y, const float* c,
                const float* z, int coeff_per_sample, int batch, size_t per_sample, void* stream) {
  DPX_REQUIRE(out && x, "null argument");
  return launch_lincomb(out, a, x, b, y, c, z, coeff_per_sample, batch, per_sample, (cudaStream_t)stream);
}

int dpx_grad_apply(const float* x, float* y, int planes, int height, int width, int axis, int adjoint, float scale,
                   void* stream) {
  DPX_REQUIRE(x && y && x != y, "null or aliased argument");
  DPX_REQUIRE(axis == 0 || axis == 1, "axis must be 0 (H) or 1 (W)");
  return launch_grad(x, y, planes, height, width, axis, adjoint != 0, scale, (cudaStream_t)stream);
}

int dpx_pad2d(const float* in, float* out, int planes, int h_in, int w_in, int h_out, int w_out, int top, int left, void* stream) {
  DPX_REQUIRE(in && out && in != out, "null or aliased argument");
  DPX_REQUIRE(planes > 0 && h_in > 0 && w_in > 0 && h_out > 0 && w_out > 0, "bad shape");
  return launch_pad2d(in, out, planes, h_in, w_in, h_out, w_out, top, left, (cudaStream_t)stream);
}

int dpx_augment(const float* in, float* out, int planes, int height, int width, int mode, void* stream) {
  DPX_REQUIRE(in && out && in != out, "null or aliased argument");
  DPX_REQUIRE(mode >= 0 && mode < 8, "mode must be 0..7");
  return launch_augment(in, out, planes, height, width, mode, (cudaStream_t)stream);
}

int dpx_axpby(float* out, float a, const float* x, float b, const float* y, size_t n, void* stream) {
  DPX_REQUIRE(out && x, "null argument");
  return launch_axpby(out, a, x, b, y, n, (cudaStream_t)stream);
}
int dpx_mul_apply(float* out, const float* x, const float* w, int w_batch, int batch, size_t per_sample, void* stream) {
  DPX_REQUIRE(out && x && w, "null argument");
  DPX_REQUIRE(w_batch == 1 || w_batch == batch, "w_batch must be 1 or batch");
  return launch_mul(out, x, w, w_batch, batch, per_sample, (cudaStream_t)stream);
}
int dpx_absmax(const float* x, float* out, int batch, size_t per_sample, void* stream) {
  DPX_REQUIRE(x && out, "null argument");
  return launch_absmax(x, out, batch, per_sample, (cudaStream_t)stream);
}

int dpx_cg_dot(const float* x, const float* y, float* dots, int batch, size_t per_sample, void* stream) {
  DPX_REQUIRE(x && y && dots, "null argument");
  return launch_cg_dot(x, y, dots, batch, per_sample, (cudaStream_t)stream);
}
int dpx_cg_update(float* x, float* r, const float* p, const float* q, const float* gamma, const float* pq,
                  float* gamma_new, int batch, size_t per_sample, void* stream) {
  DPX_REQUIRE(x && r && p && q && gamma && pq && gamma_new, "null argument");
  return launch_cg_update(x, r, p, q, gamma, pq, gamma_new, batch, per_sample, (cudaStream_t)stream);
}
int dpx_cg_direction(float* p, const float* r, const float* gamma_new, const float* gamma_old, int batch,
                     size_t per_sample, void* stream) {
  DPX_REQUIRE(p && r && gamma_new && gamma_old, "null argument");
  return launch_cg_direction(p, r, gamma_new, gamma_old, batch, per_sample, (cudaStream_t)stream);
}

int dpx_cg_gate(const float* val, const float* tol, int tol_n, int strict, float* pq, int* done, int batch, void* stream) {
  DPX_REQUIRE(val && tol && pq && done, "null argument");
  DPX_REQUIRE(batch > 0 && (tol_n == 1 || tol_n == batch), "tol_n must be 1 or batch");
  return launch_cg_gate(val, tol, tol_n, strict != 0, pq, done, batch, (cudaStream_t)stream);
}

int dpx_resid_reduce(const float* resid, float* out, int n, int batch, void* stream) {
  DPX_REQUIRE(resid && out && n > 0 && batch > 0, "bad argument");
  return launch_resid_reduce(resid, out, n, batch, (cudaStream_t)stream);
}

int dpx_solve_host(dpx_plan* p, const float* x0_host, float* x_out_host, const float* rho_host, const float* lam_host,
                   int n_iters, void* stream) {
  DPX_REQUIRE(p && x0_host && x_out_host && rho_host, "null argument");
  DPX_REQUIRE(!p->has_external, "plan has an external prox term");
  DPX_REQUIRE(n_iters > 0, "n_iters must be positive");
  cudaStream_t s = (cudaStream_t)stream;
  const Geom& g = p->g;
  const size_t nbytes = sizeof(float) * g.P * g.plane;
  const int m = p->d.n_psi;
  DPX_REQUIRE(m == 0 || lam_host, "lam_host is NULL");
  int rc;
  if (!p->hx) { rc = dev_alloc(p, (void**)&p->hx, nbytes); if (rc) return rc; }
  const bool need_v = p->d.algo != DPX_ALGO_PGD, need_u = need_v && p->d.algo != DPX_ALGO_HQS;
  for (int i = 0; i < m; ++i) {
    // a stacked-gradient term (grad2d / iso_tv) carries [B,2C,H,W] state
    const size_t nb = p->d.psi[i].linop == DPX_LINOP_GRAD_HW ? 2 * nbytes : nbytes;
    if (need_v && !p->hv[i]) { rc = dev_alloc(p, (void**)&p->hv[i], nb); if (rc) return rc; }
    if (need_u && !p->hu[i]) { rc = dev_alloc(p, (void**)&p->hu[i], nb); if (rc) return rc; }
  }
  const int nsched = n_iters * (1 + m);
  if (p->hsched_cap < nsched) {
    cudaFree(p->hsched); p->hsched = nullptr;
    rc = dev_alloc(p, (void**)&p->hsched, sizeof(float) * nsched);
    if (rc) return rc;
    p->hsched_cap = nsched;
  }
  DPX_CUDA(cudaMemcpyAsync(p->hx, x0_host, nbytes, cudaMemcpyHostToDevice, s));
  DPX_CUDA(cudaMemcpyAsync(p->hsched, rho_host, sizeof(float) * n_iters, cudaMemcpyHostToDevice, s));
  if (m) DPX_CUDA(cudaMemcpyAsync(p->hsched + n_iters, lam_host, sizeof(float) * n_iters * m, cudaMemcpyHostToDevice, s));
  const float* lam[DPX_MAX_PSI];
  int lstride[DPX_MAX_PSI];
  for (int i = 0; i < m; ++i) { lam[i] = p->hsched + n_iters * (1 + i); lstride[i] = 0; }
  rc = dpx_init_state(p, p->hx, p->hv, p->hu, stream);
  if (rc) return rc;
  rc = dpx_iters(p, p->hx, p->hv, p->hu, p->hsched, 0, lam, lstride, 0, n_iters, nullptr, stream);
  if (rc) return rc;
  DPX_CUDA(cudaMemcpyAsync(x_out_host, p->hx, nbytes, cudaMemcpyDeviceToHost, s));
  DPX_CUDA(cudaStreamSynchronize(s));
  return DPX_OK;
}

}  // extern "C"
