// dpx_kernels.cuh — launch wrappers of the element-wise / stencil / spectral kernels (dpx_kernels.cu).
#pragma once
#include "dpx_common.cuh"

namespace dpx {

// t = sum_i scale_i * A_i^T (v_i - u_i)         [ADMM: hqs=false]   /  A_i^T v_i   [HQS: hqs=true]
int launch_rhs(const Geom& g, const PsiPack& psi, bool hqs, float* t, cudaStream_t s);

// spec <- (fb + rho*spec + eps) / (dq + rho*(dpsi + wid) + eps) * inv_n     sum_square.py:150-152
int launch_spec_solve(const Geom& g, float2* spec, const float2* fb, const float* dq, int dq_batch,
                      const float* dpsi, float wid, float eps, float inv_n, RhoRef rho, cudaStream_t s);

// spec <- (spec*(1 - rho*dq) + rho*fb) * inv_n        PGD gradient step in the Fourier domain (pgd.py:39-43)
int launch_spec_pgd(const Geom& g, float2* spec, const float2* fb, const float* dq, int dq_batch, float inv_n,
                    RhoRef rho, cudaStream_t s);

// z-update + dual update (admm.py:54-57 / hqs.py:13-15); optionally fuses the next iteration's rhs
// (all-identity linops) and the residual sums {|r|^2,|s|^2,|Kx|^2,|v|^2} per sample.
int launch_prox_dual(const Geom& g, const PsiPack& psi, const float* x, bool hqs, bool skip_external, int it,
                     float* t_fused, RhoRef rho, float* resid, cudaStream_t s);

// spatial-diagonal x-update (sum_square.py:154): x = (ktb + rho*sum_i s_i b_i) / (dq + rho*wid + eps)
int launch_spatial_xupdate(const Geom& g, const PsiPack& psi, bool hqs, bool vxu, const float* ktb, const float* dq,
                           int dq_batch, const float* dpsi, float wid, float eps, bool eps_delta, RhoRef rho, float* x,
                           cudaStream_t s);
int launch_spatial_xupdate_bwd(const Geom& g, const float* gin, const float* x, const float* ktb, const float* dq, int dq_batch,
                               const float* dpsi, float wid, float eps, bool eps_delta, RhoRef rho, float* g_ktb, float* g_rho,
                               int g_rho_stride, cudaStream_t s);
// zero-pad / crop copy and the 8 dihedral image transforms (data movement of conv_doe(circular=False) and the x8 prior)
int launch_pad2d(const float* in, float* out, int planes, int hi, int wi, int ho, int wo, int top, int left, cudaStream_t s);
int launch_augment(const float* in, float* out, int planes, int h, int w, int mode, cudaStream_t s);

// ADMM_vxu (admm.py:107-120): x_i = prox(K_i z - u_i); t = sum_i s_i (x_i + u_i)   |   u_i += x_i - z
int launch_vxu_prox(const Geom& g, const PsiPack& psi, const float* z, int it, float* t, cudaStream_t s);
int launch_vxu_dual(const Geom& g, const PsiPack& psi, const float* z, cudaStream_t s);

// PGD spatial: out = x - rho*(dq*x - ktb)
int launch_pgd_spatial_step(const Geom& g, const float* x, const float* ktb, const float* dq, int dq_batch,
                            RhoRef rho, float* out, cudaStream_t s);

// v_i = K_i x0 (affine), u_i = 0                      admm.py:61-67
int launch_init_state(const Geom& g, const PsiPack& psi, const float* x, bool with_u, cudaStream_t s);

// spec <- spec * otf (or conj otf) * inv_n            conv.py:31-41
int launch_mul_otf(const Geom& g, float2* spec, const float2* otf, int otf_batch, bool conj, float inv_n,
                   cudaStream_t s);

// generic helpers
// out = prox(v, lam[b*lam_stride + it]) with the wrapper chain
int launch_prox_apply(const ProxSpec& ps, const float* v, const float* lam, int lam_stride, int it, const float* off,
                      float* out, int batch, size_t per_sample, cudaStream_t s);
int launch_lincomb(float* out, const float* a, const float* x, const float* b, const float* y, const float* c,
                   const float* z, int coeff_per_sample, int batch, size_t per_sample, cudaStream_t s);
int launch_grad(const float* x, float* y, int planes, int H, int W, int axis, bool adjoint, float scale,
                cudaStream_t s);
int launch_dual_external(const float* w, const float* v_new, float* v, float* u, size_t n, cudaStream_t s);
int launch_axpby(float* out, float a, const float* x, float b, const float* y, size_t n, cudaStream_t s);
int launch_mul(float* out, const float* x, const float* w, int w_batch, int batch, size_t per_sample, cudaStream_t s);
int launch_absmax(const float* x, float* out, int batch, size_t per_sample, cudaStream_t s);
int launch_resid_reduce(const float* resid, float* out, int n, int batch, cudaStream_t s);

// backward of the Fourier-diagonal x-update and of the native prox bodies (autograd contract)
int launch_spec_solve_bwd(const Geom& g, float2* spec, const float2* qspec, const float2* fb, const float* dq, int dq_batch,
                          const float* dpsi, float wid, float eps, float inv_n, RhoRef rho, float* g_rho, int g_rho_stride,
                          cudaStream_t s);
int launch_prox_bwd(const ProxSpec& ps, const float* v, const float* lam, int lam_stride, const float* off, const float* g,
                    float* gv, float* glam, int batch, size_t per_sample, cudaStream_t s);

// CG
int launch_cg_dot(const float* x, const float* y, float* dots, int batch, size_t per_sample, cudaStream_t s);
int launch_cg_update(float* x, float* r, const float* p, const float* q, const float* gamma, const float* pq,
                     float* gamma_new, int batch, size_t per_sample, cudaStream_t s);
int launch_cg_direction(float* p, const float* r, const float* gn, const float* go, int batch, size_t per_sample,
                        cudaStream_t s);
int launch_cg_gate(const float* val, const float* tol, int tol_n, bool strict, float* pq, int* done, int batch, cudaStream_t s);

}  // namespace dpx
