// dpx_kernels.cu — element-wise, stencil and spectral-diagonal kernels of the proximal iteration.
//
// All kernels are HBM-bound streaming kernels: one 128-bit vector per thread per operand, planes on
// gridDim.y (so the sample index, its rho/lam and the residual slot are block-uniform), no shared
// memory except for the block reductions of the residual norms.
//
// Reference arithmetic being replaced (paths relative to /root/reference):
//   k_rhs           : `b_i = v_i - u_i`, `Ktb += rho * K_i^T b_i`       algo/admm.py:51, proxfn/sum_square.py:133-134
//   k_spec_solve    : `(fftn(Ktb)+eps)/(diag+eps)`                      proxfn/sum_square.py:142-152
//   k_prox_dual     : `v_i = prox_i(Kx_i+u_i)`, `u_i += Kx_i - v_i`     algo/admm.py:54-57, proxfn/base.py:55-64
//   k_spatial_x     : `Ktb/(diag+eps)`                                  proxfn/sum_square.py:154
//   k_vxu_*         : ADMM_vxu ordering                                 algo/admm.py:107-120
//   k_spec_pgd      : `x - rho * K^T(Kx - b)`                           algo/pgd.py:39-43, proxfn/sum_square.py:29-32
//   k_cg_*          : cg's fused vector updates                         linalg/solve/solver_cg.py:109-129
#include "dpx_kernels.cuh"

namespace dpx {

namespace {

constexpr int kThreads = 256;

template <int VEC>
__device__ __forceinline__ void loadv(float (&r)[VEC], const float* p) {
  if constexpr (VEC == 4) {
    const float4 t = ld4(p);
    r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
  } else {
    r[0] = *p;
  }
}
template <int VEC>
__device__ __forceinline__ void storev(float* p, const float (&r)[VEC]) {
  if constexpr (VEC == 4) {
    st4(p, make_float4(r[0], r[1], r[2], r[3]));
  } else {
    *p = r[0];
  }
}

// A_i applied at (h, w0..w0+VEC) of one plane `xp` (circular forward differences, linop/grad.py:8-23).
template <int VEC>
__device__ __forceinline__ void apply_linop(int linop, float scale, const float* __restrict__ xp, const float (&xv)[VEC],
                                            int h, int w0, int H, int W, float (&out)[VEC]) {
  if (linop == DPX_LINOP_IDENTITY) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) out[k] = scale * xv[k];
  } else if (linop == DPX_LINOP_GRAD_H) {
    const int hn = (h + 1 == H) ? 0 : h + 1;
    float xn[VEC];
    loadv<VEC>(xn, xp + (size_t)hn * W + w0);
#pragma unroll
    for (int k = 0; k < VEC; ++k) out[k] = scale * (xn[k] - xv[k]);
  } else {  // GRAD_W
    const int wn = (w0 + VEC == W) ? 0 : w0 + VEC;
    const float last = xp[(size_t)h * W + wn];
#pragma unroll
    for (int k = 0; k < VEC; ++k) out[k] = scale * ((k + 1 < VEC ? xv[k + 1] : last) - xv[k]);
  }
}

// scale * A_i^T d at (h, w0..): identity -> d ; grad -> d[prev] - d[cur]
template <int VEC, typename LoadD>
__device__ __forceinline__ void apply_adjoint(int linop, float scale, LoadD&& loadd, int h, int w0, int H, int W,
                                              float (&out)[VEC]) {
  float d[VEC];
  loadd(h, w0, d);
  if (linop == DPX_LINOP_IDENTITY) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) out[k] = scale * d[k];
  } else if (linop == DPX_LINOP_GRAD_H) {
    const int hp = (h == 0) ? H - 1 : h - 1;
    float dp[VEC];
    loadd(hp, w0, dp);
#pragma unroll
    for (int k = 0; k < VEC; ++k) out[k] = scale * (dp[k] - d[k]);
  } else {
    const int wp = (w0 == 0) ? W - 1 : w0 - 1;
    float prev[1];
    // scalar fetch of the element left of the vector
    float tmp[VEC];
    if constexpr (VEC == 1) {
      loadd(h, wp, tmp);
      prev[0] = tmp[0];
    } else {
      const int wv = wp - (wp % VEC);
      loadd(h, wv, tmp);
      prev[0] = tmp[wp - wv];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) out[k] = scale * ((k == 0 ? prev[0] : d[k - 1]) - d[k]);
  }
}

__device__ __forceinline__ float rho_of(const RhoRef& r, int b) { return r.p[(size_t)b * r.stride + r.it]; }

// ------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kThreads) k_rhs(Geom g, PsiPack psi, bool hqs, float* __restrict__ t) {
  const int p = blockIdx.y;
  const size_t vi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t e = vi * VEC;
  if (e >= g.plane) return;
  const int h = (int)(e / g.W), w0 = (int)(e - (size_t)h * g.W);
  const size_t base = (size_t)p * g.plane;
  float acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
  for (int i = 0; i < psi.n; ++i) {
    const PsiTerm& tm = psi.t[i];
    if (tm.linop == DPX_LINOP_GRAD_HW) {            // stacked [grad_H; grad_W] term: state is [B,2C,H,W]
      const int b = p / g.C, c = p - b * g.C;
#pragma unroll
      for (int comp = 0; comp < 2; ++comp) {
        const size_t cb = ((size_t)b * 2 * g.C + (size_t)comp * g.C + c) * g.plane;
        const float* vp2 = tm.v + cb;
        const float* up2 = hqs ? nullptr : tm.u + cb;
        auto loadd2 = [&](int hh, int ww, float(&d)[VEC]) {
          loadv<VEC>(d, vp2 + (size_t)hh * g.W + ww);
          if (up2) {
            float uu[VEC];
            loadv<VEC>(uu, up2 + (size_t)hh * g.W + ww);
#pragma unroll
            for (int k = 0; k < VEC; ++k) d[k] -= uu[k];
          }
        };
        float o[VEC];
        apply_adjoint<VEC>(comp == 0 ? DPX_LINOP_GRAD_H : DPX_LINOP_GRAD_W, tm.scale, loadd2, h, w0, g.H, g.W, o);
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] += o[k];
      }
      continue;
    }
    const float* vp = tm.v + base;
    const float* up = hqs ? nullptr : tm.u + base;
    auto loadd = [&](int hh, int ww, float(&d)[VEC]) {
      loadv<VEC>(d, vp + (size_t)hh * g.W + ww);
      if (up) {
        float uu[VEC];
        loadv<VEC>(uu, up + (size_t)hh * g.W + ww);
#pragma unroll
        for (int k = 0; k < VEC; ++k) d[k] -= uu[k];
      }
    };
    float o[VEC];
    apply_adjoint<VEC>(tm.linop, tm.scale, loadd, h, w0, g.H, g.W, o);
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] += o[k];
  }
  storev<VEC>(t + base + e, acc);
}

// ------------------------------------------------------------------------------------------------
template <int CV>  // complex elements per thread (1 or 2)
__global__ void __launch_bounds__(kThreads)
    k_spec_solve(Geom g, float2* __restrict__ spec, const float2* __restrict__ fb, const float* __restrict__ dq,
                 int dq_batch, const float* __restrict__ dpsi, float wid, float eps, float inv_n, RhoRef rho) {
  const int p = blockIdx.y;
  const size_t ci = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * CV;
  if (ci >= g.splane) return;
  const int b = p / g.C, c = p - b * g.C;
  const float r = rho_of(rho, b);
  const size_t off = (size_t)p * g.splane + ci;
  const size_t doff = (size_t)(dq_batch > 1 ? p : c) * g.splane + ci;
  const size_t poff = (size_t)c * g.splane + ci;
  float2 s[CV], f[CV];
  float q[CV], ps[CV];
  if constexpr (CV == 2) {
    const float4 sv = *reinterpret_cast<const float4*>(spec + off);
    s[0] = make_float2(sv.x, sv.y); s[1] = make_float2(sv.z, sv.w);
    if (fb) {
      const float4 fv = *reinterpret_cast<const float4*>(fb + off);
      f[0] = make_float2(fv.x, fv.y); f[1] = make_float2(fv.z, fv.w);
    } else {
      f[0] = f[1] = make_float2(0.f, 0.f);
    }
    if (dq) { const float2 t = *reinterpret_cast<const float2*>(dq + doff); q[0] = t.x; q[1] = t.y; } else { q[0] = q[1] = 0.f; }
    if (dpsi) { const float2 t = *reinterpret_cast<const float2*>(dpsi + poff); ps[0] = t.x; ps[1] = t.y; } else { ps[0] = ps[1] = 0.f; }
  } else {
    s[0] = spec[off];
    f[0] = fb ? fb[off] : make_float2(0.f, 0.f);
    q[0] = dq ? dq[doff] : 0.f;
    ps[0] = dpsi ? dpsi[poff] : 0.f;
  }
#pragma unroll
  for (int k = 0; k < CV; ++k) {
    const float den = q[k] + r * (ps[k] + wid) + eps;
    const float re = (f[k].x + r * s[k].x + eps) / den;
    const float im = (f[k].y + r * s[k].y) / den;
    s[k] = make_float2(re * inv_n, im * inv_n);
  }
  if constexpr (CV == 2) {
    *reinterpret_cast<float4*>(spec + off) = make_float4(s[0].x, s[0].y, s[1].x, s[1].y);
  } else {
    spec[off] = s[0];
  }
}

__global__ void __launch_bounds__(kThreads)
    k_spec_pgd(Geom g, float2* __restrict__ spec, const float2* __restrict__ fb, const float* __restrict__ dq, int dq_batch,
               float inv_n, RhoRef rho) {
  const int p = blockIdx.y;
  const size_t ci = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ci >= g.splane) return;
  const int b = p / g.C, c = p - b * g.C;
  const float r = rho_of(rho, b);
  const size_t off = (size_t)p * g.splane + ci;
  const float q = dq ? dq[(size_t)(dq_batch > 1 ? p : c) * g.splane + ci] : 0.f;
  const float2 f = fb ? fb[off] : make_float2(0.f, 0.f);
  float2 s = spec[off];
  // x - rho*(|O|^2 X - conj(O) B)  in the Fourier domain
  s.x = (s.x - r * (q * s.x - f.x)) * inv_n;
  s.y = (s.y - r * (q * s.y - f.y)) * inv_n;
  spec[off] = s;
}

__global__ void __launch_bounds__(kThreads)
    k_mul_otf(Geom g, float2* __restrict__ spec, const float2* __restrict__ otf, int otf_batch, bool conj, float inv_n) {
  const int p = blockIdx.y;
  const size_t ci = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ci >= g.splane) return;
  const int c = p % g.C;
  const float2 o = otf[(size_t)(otf_batch > 1 ? p : c) * g.splane + ci];
  const float oy = conj ? -o.y : o.y;
  const size_t off = (size_t)p * g.splane + ci;
  const float2 s = spec[off];
  spec[off] = make_float2((s.x * o.x - s.y * oy) * inv_n, (s.x * oy + s.y * o.x) * inv_n);
}

// ------------------------------------------------------------------------------------------------
template <int VEC, bool FUSE, bool RESID>
__global__ void __launch_bounds__(kThreads)
    k_prox_dual(Geom g, PsiPack psi, const float* __restrict__ x, bool hqs, bool skip_external, int it,
                float* __restrict__ t, RhoRef rho, float* __restrict__ resid) {
  __shared__ float red[4 * 32];
  const int p = blockIdx.y;
  const int b = p / g.C;
  const size_t vi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t e = vi * VEC;
  const bool active = e < g.plane;
  float racc[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
    const int h = (int)(e / g.W), w0 = (int)(e - (size_t)h * g.W);
    const size_t base = (size_t)p * g.plane;
    const float* xp = x + base;
    float xv[VEC];
    loadv<VEC>(xv, xp + e);
    float tacc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) tacc[k] = 0.f;
    const float r = RESID ? rho_of(rho, b) : 0.f;
    for (int i = 0; i < psi.n; ++i) {
      const PsiTerm& tm = psi.t[i];
      float kx[VEC], w[VEC], offv[VEC], vn[VEC], un[VEC];
      if (tm.linop == DPX_LINOP_GRAD_HW) {
        // K x = [grad_H x ; grad_W x] stacked on the channel axis; DPX_PROX_ISO_TV couples the two components
        // (group soft-threshold, v = max(1 - lam/|w|_2, 0) w), any other native prox acts element-wise on both.
        const int c = p - b * g.C;
        const size_t cb0 = ((size_t)b * 2 * g.C + c) * g.plane + e, cb1 = cb0 + (size_t)g.C * g.plane;
        float kh[VEC], kw[VEC], wh[VEC], ww[VEC], vh[VEC], vw[VEC];
        apply_linop<VEC>(DPX_LINOP_GRAD_H, tm.scale, xp, xv, h, w0, g.H, g.W, kh);
        apply_linop<VEC>(DPX_LINOP_GRAD_W, tm.scale, xp, xv, h, w0, g.H, g.W, kw);
        if (!hqs) {
          float u0[VEC], u1[VEC];
          loadv<VEC>(u0, tm.u + cb0); loadv<VEC>(u1, tm.u + cb1);
#pragma unroll
          for (int k = 0; k < VEC; ++k) { wh[k] = kh[k] + u0[k]; ww[k] = kw[k] + u1[k]; }
        } else {
#pragma unroll
          for (int k = 0; k < VEC; ++k) { wh[k] = kh[k]; ww[k] = kw[k]; }
        }
        const float lam = tm.lam[(size_t)b * tm.lam_stride + it];
        const ProxSpec ps{tm.prox, tm.alpha, tm.beta, tm.inv_beta, tm.lo, tm.hi};
        float v0o[VEC], v1o[VEC];
        if (RESID) { loadv<VEC>(v0o, tm.v + cb0); loadv<VEC>(v1o, tm.v + cb1); }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          if (tm.prox == DPX_PROX_ISO_TV) {
            const float lam_eff = tm.beta * tm.beta * lam * tm.alpha;
            const float a0 = tm.beta * wh[k], a1 = tm.beta * ww[k];
            const float nrm = sqrtf(a0 * a0 + a1 * a1);
            const float f = nrm > lam_eff ? (1.f - lam_eff / nrm) * tm.inv_beta : 0.f;
            vh[k] = f * a0; vw[k] = f * a1;
          } else {
            vh[k] = prox_wrapped(ps, wh[k], lam, 0.f); vw[k] = prox_wrapped(ps, ww[k], lam, 0.f);
          }
        }
        storev<VEC>(tm.v + cb0, vh); storev<VEC>(tm.v + cb1, vw);
        if (!hqs) {
          float u0[VEC], u1[VEC];
#pragma unroll
          for (int k = 0; k < VEC; ++k) { u0[k] = wh[k] - vh[k]; u1[k] = ww[k] - vw[k]; }
          storev<VEC>(tm.u + cb0, u0); storev<VEC>(tm.u + cb1, u1);
        }
        if (RESID) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            const float r0 = kh[k] - vh[k], r1 = kw[k] - vw[k];
            const float s0 = r * tm.scale * (vh[k] - v0o[k]), s1 = r * tm.scale * (vw[k] - v1o[k]);
            racc[0] += r0 * r0 + r1 * r1; racc[1] += s0 * s0 + s1 * s1;
            racc[2] += kh[k] * kh[k] + kw[k] * kw[k]; racc[3] += vh[k] * vh[k] + vw[k] * vw[k];
          }
        }
        continue;
      }
      apply_linop<VEC>(tm.linop, tm.scale, xp, xv, h, w0, g.H, g.W, kx);
      if (tm.off) {
        loadv<VEC>(offv, tm.off + base + e);
#pragma unroll
        for (int k = 0; k < VEC; ++k) kx[k] -= offv[k];
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) offv[k] = 0.f;
      }
      if (!hqs) {
        float uo[VEC];
        loadv<VEC>(uo, tm.u + base + e);
#pragma unroll
        for (int k = 0; k < VEC; ++k) w[k] = kx[k] + uo[k];
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) w[k] = kx[k];
      }
      if (tm.prox == DPX_PROX_EXTERNAL) {
        if (skip_external) storev<VEC>(tm.v + base + e, w);   // caller evaluates its prox on w
        continue;
      }
      const float lam = tm.lam[(size_t)b * tm.lam_stride + it];
      const ProxSpec ps{tm.prox, tm.alpha, tm.beta, tm.inv_beta, tm.lo, tm.hi};
      float vo[VEC];
      if (RESID) loadv<VEC>(vo, tm.v + base + e);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        vn[k] = prox_wrapped(ps, w[k], lam, offv[k]);
        un[k] = w[k] - vn[k];
      }
      storev<VEC>(tm.v + base + e, vn);
      if (!hqs) storev<VEC>(tm.u + base + e, un);
      if (FUSE) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) tacc[k] += tm.scale * (hqs ? vn[k] : vn[k] - un[k]);
      }
      if (RESID) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float rr = kx[k] - vn[k], ss = r * tm.scale * (vn[k] - vo[k]);
          racc[0] += rr * rr; racc[1] += ss * ss; racc[2] += kx[k] * kx[k]; racc[3] += vn[k] * vn[k];
        }
      }
    }
    if (FUSE) storev<VEC>(t + base + e, tacc);
  }
  if (RESID) {
    block_sum<4>(racc, red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(resid + (size_t)b * 4 + k, racc[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_spatial_x(Geom g, PsiPack psi, bool hqs, bool vxu, const float* __restrict__ ktb, const float* __restrict__ dq,
                int dq_batch, const float* __restrict__ dpsi, float wid, float eps, bool eps_delta, RhoRef rho,
                float* __restrict__ x) {
  const int p = blockIdx.y;
  const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= g.plane) return;
  const int b = p / g.C, c = p - b * g.C;
  const float r = rho_of(rho, b);
  const size_t base = (size_t)p * g.plane;
  float num[VEC], den[VEC];
  if (ktb) loadv<VEC>(num, ktb + base + e);
  else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) num[k] = 0.f;
  }
  if (dq) loadv<VEC>(den, dq + (size_t)(dq_batch > 1 ? p : c) * g.plane + e);
  else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) den[k] = 0.f;
  }
  for (int i = 0; i < psi.n; ++i) {
    const PsiTerm& tm = psi.t[i];
    float vv[VEC], uu[VEC];
    loadv<VEC>(vv, tm.v + base + e);
    if (!hqs) {
      loadv<VEC>(uu, tm.u + base + e);
#pragma unroll
      for (int k = 0; k < VEC; ++k) vv[k] = vxu ? vv[k] + uu[k] : vv[k] - uu[k];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) num[k] += r * (tm.scale * vv[k]);
  }
  if (eps_delta && e == 0) num[0] += eps;
  if (dpsi) {                // mask-type psi linops: the denominator gains rho * sum_i s_i^2 diag_i (sum_square.py:145-148)
    float dp[VEC];
    loadv<VEC>(dp, dpsi + (size_t)c * g.plane + e);
#pragma unroll
    for (int k = 0; k < VEC; ++k) den[k] += r * dp[k];
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) num[k] = num[k] / (den[k] + r * wid + eps);
  storev<VEC>(x + base + e, num);
}

// Backward of the spatial-diagonal x-update x = (ktb + rho t + eps delta_0) / D, D = dq + rho (dpsi + wid) + eps:
//   g_ktb = g / D   (dL/dt = rho g_ktb),   g_rho[b] = sum g_ktb (x (dq + eps) - ktb - eps delta_0) / rho
template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_spatial_x_bwd(Geom g, const float* __restrict__ gin, const float* __restrict__ x, const float* __restrict__ ktb,
                    const float* __restrict__ dq, int dq_batch, const float* __restrict__ dpsi, float wid, float eps,
                    bool eps_delta, RhoRef rho, float* __restrict__ g_ktb, float* __restrict__ g_rho, int g_rho_stride) {
  __shared__ float red[32];
  const int p = blockIdx.y;
  const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  const int b = p / g.C, c = p - b * g.C;
  const float r = rho_of(rho, b);
  const size_t base = (size_t)p * g.plane;
  float acc[1] = {0.f};
  if (e < g.plane) {
    float gv[VEC], d0[VEC], dp[VEC], kb[VEC], xv[VEC];
    loadv<VEC>(gv, gin + base + e);
    if (dq) loadv<VEC>(d0, dq + (size_t)(dq_batch > 1 ? p : c) * g.plane + e);
    if (dpsi) loadv<VEC>(dp, dpsi + (size_t)c * g.plane + e);
    if (g_rho) {
      loadv<VEC>(xv, x + base + e);
      if (ktb) loadv<VEC>(kb, ktb + base + e);
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float q = dq ? d0[k] : 0.f;
      const float D = q + r * ((dpsi ? dp[k] : 0.f) + wid) + eps;
      gv[k] = gv[k] / D;
      if (g_rho) {
        float num0 = ktb ? kb[k] : 0.f;
        if (eps_delta && e + k == 0) num0 += eps;
        acc[0] += gv[k] * (xv[k] * (q + eps) - num0) / r;
      }
    }
    storev<VEC>(g_ktb + base + e, gv);
  }
  if (g_rho) {
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) atomicAdd(g_rho + (size_t)b * g_rho_stride, acc[0]);
  }
}

// out[y][x] = in[y - top][x - left] where that lies inside the input, else 0: zero padding (top, left >= 0) and cropping
// (negative offsets) of the `circular=False` convolutions (linop/conv.py:100-121, contrib/optic/common.py:97-117)
__global__ void __launch_bounds__(kThreads)
    k_pad2d(const float* __restrict__ in, float* __restrict__ out, int hi, int wi, int ho, int wo, int top, int left) {
  const int p = blockIdx.y;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)ho * wo) return;
  const int y = (int)(e / wo), xx = (int)(e - (size_t)y * wo);
  const int sy = y - top, sx = xx - left;
  out[(size_t)p * ho * wo + e] = (sy >= 0 && sy < hi && sx >= 0 && sx < wi) ? in[((size_t)p * hi + sy) * wi + sx] : 0.f;
}

// the 8 flips / rotations of Augment.augment (pnp/denoisers/composite.py:30-47); out is [wo = h][..] for the transposing modes
__global__ void __launch_bounds__(kThreads)
    k_augment(const float* __restrict__ in, float* __restrict__ out, int h, int w, int mode) {
  const int p = blockIdx.y;
  const bool tr = mode == 1 || mode == 3 || mode == 5 || mode == 7;       // output is [w, h]
  const int ho = tr ? w : h, wo = tr ? h : w;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)h * w) return;
  const int i = (int)(e / wo), j = (int)(e - (size_t)i * wo);
  int sy, sx;                                                             // source pixel of output (i, j)
  switch (mode) {
    case 0: sy = i; sx = j; break;
    case 1: sy = j; sx = i; break;                      // rot90(1).flip(rows)   = transpose
    case 2: sy = ho - 1 - i; sx = j; break;             // flip(rows)
    case 3: sy = h - 1 - j; sx = i; break;              // rot90(3)
    case 4: sy = i; sx = wo - 1 - j; break;             // rot90(2).flip(rows)   = flip(cols)
    case 5: sy = j; sx = w - 1 - i; break;              // rot90(1)
    case 6: sy = ho - 1 - i; sx = wo - 1 - j; break;    // rot90(2)
    default: sy = h - 1 - j; sx = w - 1 - i; break;     // rot90(3).flip(rows)   = anti-transpose
  }
  out[(size_t)p * h * w + e] = in[((size_t)p * h + sy) * w + sx];
}

// ------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_vxu_prox(Geom g, PsiPack psi, const float* __restrict__ z, int it, float* __restrict__ t) {
  const int p = blockIdx.y;
  const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= g.plane) return;
  const int b = p / g.C;
  const size_t base = (size_t)p * g.plane;
  float zv[VEC], tacc[VEC];
  loadv<VEC>(zv, z + base + e);
#pragma unroll
  for (int k = 0; k < VEC; ++k) tacc[k] = 0.f;
  for (int i = 0; i < psi.n; ++i) {
    const PsiTerm& tm = psi.t[i];
    float uo[VEC], offv[VEC], xn[VEC];
    loadv<VEC>(uo, tm.u + base + e);
    if (tm.off) loadv<VEC>(offv, tm.off + base + e);
    else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) offv[k] = 0.f;
    }
    const float lam = tm.lam[(size_t)b * tm.lam_stride + it];
    const ProxSpec ps{tm.prox, tm.alpha, tm.beta, tm.inv_beta, tm.lo, tm.hi};
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float kz = tm.scale * zv[k] - offv[k];
      xn[k] = prox_wrapped(ps, kz - uo[k], lam, offv[k]);
      tacc[k] += tm.scale * (xn[k] + uo[k]);
    }
    storev<VEC>(tm.v + base + e, xn);
  }
  if (t) storev<VEC>(t + base + e, tacc);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) k_vxu_dual(Geom g, PsiPack psi, const float* __restrict__ z) {
  const int p = blockIdx.y;
  const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= g.plane) return;
  const size_t base = (size_t)p * g.plane;
  float zv[VEC];
  loadv<VEC>(zv, z + base + e);
  for (int i = 0; i < psi.n; ++i) {
    const PsiTerm& tm = psi.t[i];
    float uo[VEC], xi[VEC];
    loadv<VEC>(uo, tm.u + base + e);
    loadv<VEC>(xi, tm.v + base + e);
#pragma unroll
    for (int k = 0; k < VEC; ++k) uo[k] = uo[k] + xi[k] - zv[k];
    storev<VEC>(tm.u + base + e, uo);
  }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_pgd_spatial(Geom g, const float* __restrict__ x, const float* __restrict__ ktb, const float* __restrict__ dq,
                  int dq_batch, RhoRef rho, float* __restrict__ out) {
  const int p = blockIdx.y;
  const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= g.plane) return;
  const int b = p / g.C, c = p - b * g.C;
  const float r = rho_of(rho, b);
  const size_t base = (size_t)p * g.plane;
  float xv[VEC], kb[VEC], q[VEC];
  loadv<VEC>(xv, x + base + e);
  loadv<VEC>(kb, ktb + base + e);
  loadv<VEC>(q, dq + (size_t)(dq_batch > 1 ? p : c) * g.plane + e);
#pragma unroll
  for (int k = 0; k < VEC; ++k) xv[k] = xv[k] - r * (q[k] * xv[k] - kb[k]);
  storev<VEC>(out + base + e, xv);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) k_init_state(Geom g, PsiPack psi, const float* __restrict__ x, bool with_u) {
  const int p = blockIdx.y;
  const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= g.plane) return;
  const int h = (int)(e / g.W), w0 = (int)(e - (size_t)h * g.W);
  const size_t base = (size_t)p * g.plane;
  const float* xp = x + base;
  float xv[VEC], zero[VEC];
  loadv<VEC>(xv, xp + e);
#pragma unroll
  for (int k = 0; k < VEC; ++k) zero[k] = 0.f;
  for (int i = 0; i < psi.n; ++i) {
    const PsiTerm& tm = psi.t[i];
    float kx[VEC];
    if (tm.linop == DPX_LINOP_GRAD_HW) {
      const int b = p / g.C, c = p - b * g.C;
      const size_t cb0 = ((size_t)b * 2 * g.C + c) * g.plane + e, cb1 = cb0 + (size_t)g.C * g.plane;
      apply_linop<VEC>(DPX_LINOP_GRAD_H, tm.scale, xp, xv, h, w0, g.H, g.W, kx);
      storev<VEC>(tm.v + cb0, kx);
      apply_linop<VEC>(DPX_LINOP_GRAD_W, tm.scale, xp, xv, h, w0, g.H, g.W, kx);
      storev<VEC>(tm.v + cb1, kx);
      if (with_u) { storev<VEC>(tm.u + cb0, zero); storev<VEC>(tm.u + cb1, zero); }
      continue;
    }
    apply_linop<VEC>(tm.linop, tm.scale, xp, xv, h, w0, g.H, g.W, kx);
    if (tm.off) {
      float offv[VEC];
      loadv<VEC>(offv, tm.off + base + e);
#pragma unroll
      for (int k = 0; k < VEC; ++k) kx[k] -= offv[k];
    }
    storev<VEC>(tm.v + base + e, kx);
    if (with_u) storev<VEC>(tm.u + base + e, zero);
  }
}

// ------------------------------------------------------------------------------------------------
// generic helpers (flat [batch, per_sample] views; blockIdx.y = sample)
template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_prox_apply(ProxSpec ps, const float* __restrict__ v, const float* __restrict__ lam, int lam_stride, int it,
                 const float* __restrict__ off, float* __restrict__ out, size_t per_sample) {
  const int b = blockIdx.y;
  const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= per_sample) return;
  const size_t base = (size_t)b * per_sample + e;
  const float l = lam[(size_t)b * lam_stride + it];
  float vv[VEC], ov[VEC];
  loadv<VEC>(vv, v + base);
  if (off) loadv<VEC>(ov, off + base);
  else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) ov[k] = 0.f;
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) vv[k] = prox_wrapped(ps, vv[k], l, ov[k]);
  storev<VEC>(out + base, vv);
}

// isotropic (group) shrink of a stacked [B,2C,H,W] tensor: elements e and e + half of a sample form one group
template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_prox_iso(ProxSpec ps, const float* __restrict__ v, const float* __restrict__ lam, int lam_stride, int it,
               float* __restrict__ out, size_t half) {
  const int b = blockIdx.y;
  const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= half) return;
  const size_t base = (size_t)b * 2 * half + e;
  const float lam_eff = ps.beta * ps.beta * lam[(size_t)b * lam_stride + it] * ps.alpha;
  float a0[VEC], a1[VEC];
  loadv<VEC>(a0, v + base);
  loadv<VEC>(a1, v + base + half);
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const float x0 = ps.beta * a0[k], x1 = ps.beta * a1[k];
    const float nrm = sqrtf(x0 * x0 + x1 * x1);
    const float f = nrm > lam_eff ? (1.f - lam_eff / nrm) * ps.inv_beta : 0.f;
    a0[k] = f * x0; a1[k] = f * x1;
  }
  storev<VEC>(out + base, a0);
  storev<VEC>(out + base + half, a1);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_lincomb(float* __restrict__ out, const float* a, const float* __restrict__ x, const float* b,
              const float* __restrict__ y, const float* c, const float* __restrict__ z, int cps, size_t per_sample) {
  const int s = blockIdx.y;
  const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= per_sample) return;
  const size_t base = (size_t)s * per_sample + e;
  const float ca = a ? a[cps ? s : 0] : 1.f, cb = b ? b[cps ? s : 0] : 1.f, cc = c ? c[cps ? s : 0] : 1.f;
  float xv[VEC], yv[VEC], zv[VEC];
  loadv<VEC>(xv, x + base);
#pragma unroll
  for (int k = 0; k < VEC; ++k) xv[k] = ca * xv[k];
  if (y) {
    loadv<VEC>(yv, y + base);
#pragma unroll
    for (int k = 0; k < VEC; ++k) xv[k] += cb * yv[k];
  }
  if (z) {
    loadv<VEC>(zv, z + base);
#pragma unroll
    for (int k = 0; k < VEC; ++k) xv[k] += cc * zv[k];
  }
  storev<VEC>(out + base, xv);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_grad(const float* __restrict__ x, float* __restrict__ y, int H, int W, int linop, bool adjoint, float scale) {
  const int p = blockIdx.y;
  const size_t plane = (size_t)H * W;
  const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= plane) return;
  const int h = (int)(e / W), w0 = (int)(e - (size_t)h * W);
  const float* xp = x + (size_t)p * plane;
  float o[VEC];
  if (!adjoint) {
    float xv[VEC];
    loadv<VEC>(xv, xp + e);
    apply_linop<VEC>(linop, scale, xp, xv, h, w0, H, W, o);
  } else {
    auto loadd = [&](int hh, int ww, float(&d)[VEC]) { loadv<VEC>(d, xp + (size_t)hh * W + ww); };
    apply_adjoint<VEC>(linop, scale, loadd, h, w0, H, W, o);
  }
  storev<VEC>(y + (size_t)p * plane + e, o);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_dual_external(const float* __restrict__ w, const float* __restrict__ vnew, float* __restrict__ v,
                    float* __restrict__ u, size_t n) {
  const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= n) return;
  float wv[VEC], vn[VEC];
  loadv<VEC>(wv, w + e);
  loadv<VEC>(vn, vnew + e);
#pragma unroll
  for (int k = 0; k < VEC; ++k) wv[k] = wv[k] - vn[k];
  if (u) storev<VEC>(u + e, wv);
  storev<VEC>(v + e, vn);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_axpby(float* __restrict__ out, float a, const float* __restrict__ x, float b, const float* __restrict__ y, size_t n) {
  const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= n) return;
  float xv[VEC], yv[VEC];
  loadv<VEC>(xv, x + e);
  if (y) {
    loadv<VEC>(yv, y + e);
#pragma unroll
    for (int k = 0; k < VEC; ++k) xv[k] = a * xv[k] + b * yv[k];
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) xv[k] = a * xv[k];
  }
  storev<VEC>(out + e, xv);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_mul(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ w, int w_batch, size_t per_sample) {
  const int b = blockIdx.y;
  const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= per_sample) return;
  float xv[VEC], wv[VEC];
  loadv<VEC>(xv, x + (size_t)b * per_sample + e);
  loadv<VEC>(wv, w + (size_t)(w_batch > 1 ? b : 0) * per_sample + e);
#pragma unroll
  for (int k = 0; k < VEC; ++k) xv[k] = wv[k] * xv[k];
  storev<VEC>(out + (size_t)b * per_sample + e, xv);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) k_absmax(const float* __restrict__ x, float* __restrict__ out, size_t per_sample) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  float m = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x * VEC;
  for (size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; e < per_sample; e += stride) {
    float xv[VEC];
    loadv<VEC>(xv, x + (size_t)b * per_sample + e);
#pragma unroll
    for (int k = 0; k < VEC; ++k) m = fmaxf(m, fabsf(xv[k]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (warp == 0) {
    m = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    // non-negative floats order like their bit patterns
    if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(out + b), __float_as_uint(m));
  }
}

__global__ void k_resid_reduce(const float* __restrict__ resid, float* __restrict__ out, int n, int batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 4) return;
  const int it = i >> 2, k = i & 3;
  float s = 0.f;
  for (int b = 0; b < batch; ++b) s += resid[((size_t)it * batch + b) * 4 + k];
  out[i] = s;
}

// ---- CG -------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
//  Backward kernels (autograd contract of Algorithm.iter/solve).  The reference gets these
//  by plain autograd through solve_direct (sum_square.py:123-156) and the prox bodies; here they are closed forms.
// ------------------------------------------------------------------------------------------------
// spec holds W = F(g) (unnormalised R2C of the upstream gradient), qspec holds Q = F(x) of the forward output.
//   spec    <- W / Dn * inv_n                        (C2R gives dL/d(K^T b);  dL/dt = rho * that)
//   g_rho_b += sum_k wgt_k Re( conj(W_k) (Q_k (dq_k + eps) - fb_k - eps) ) / (Dn_k * n * rho_b)
// with Dn = dq + rho (dpsi + wid) + eps and wgt = 2 for the half-spectrum columns that stand for two bins.
__global__ void __launch_bounds__(kThreads)
    k_spec_solve_bwd(Geom g, float2* __restrict__ spec, const float2* __restrict__ qspec, const float2* __restrict__ fb,
                     const float* __restrict__ dq, int dq_batch, const float* __restrict__ dpsi, float wid, float eps,
                     float inv_n, RhoRef rho, float* __restrict__ g_rho, int g_rho_stride) {
  __shared__ float red[32];
  const int p = blockIdx.y;
  const int b = p / g.C, c = p - b * g.C;
  const float r = rho_of(rho, b);
  float acc[1] = {0.f};
  const size_t ci = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ci < g.splane) {
    const size_t off = (size_t)p * g.splane + ci;
    const float q = dq ? dq[(size_t)(dq_batch > 1 ? p : c) * g.splane + ci] : 0.f;
    const float ps = dpsi ? dpsi[(size_t)c * g.splane + ci] : 0.f;
    const float inv = 1.0f / (q + r * (ps + wid) + eps);
    const float2 w = spec[off];
    if (g_rho) {
      const float2 Q = qspec[off];
      const float2 f = fb ? fb[off] : make_float2(0.f, 0.f);
      const int k = (int)(ci % g.Wc);
      const float wgt = (k == 0 || (g.W % 2 == 0 && k == g.Wc - 1)) ? 1.f : 2.f;
      const float nr = Q.x * (q + eps) - f.x - eps, ni = Q.y * (q + eps) - f.y;
      acc[0] = wgt * (w.x * nr + w.y * ni) * inv * inv_n / r;
    }
    spec[off] = make_float2(w.x * inv * inv_n, w.y * inv * inv_n);
  }
  if (g_rho) {
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) atomicAdd(g_rho + (size_t)b * g_rho_stride, acc[0]);
  }
}

// backward of ProxFn.prox (wrapper chain of proxfn/base.py:55-64 around the native `_prox` bodies):
//   out = 1/beta P(beta (v - off), beta^2 alpha lam) + off   =>   g_v = P'_z g,   g_lam = beta alpha sum P'_lam g
template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_prox_bwd(ProxSpec ps, const float* __restrict__ v, const float* __restrict__ lam, int lam_stride,
               const float* __restrict__ off, const float* __restrict__ g, float* __restrict__ gv,
               float* __restrict__ glam, size_t per_sample) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  const float le = ps.beta * ps.beta * lam[(size_t)b * lam_stride] * ps.alpha;
  float acc[1] = {0.f};
  const size_t stride = (size_t)gridDim.x * blockDim.x * VEC;
  for (size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; e < per_sample; e += stride) {
    const size_t base = (size_t)b * per_sample + e;
    float vv[VEC], ov[VEC], gg[VEC];
    loadv<VEC>(vv, v + base);
    loadv<VEC>(gg, g + base);
    if (off) loadv<VEC>(ov, off + base);
    else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) ov[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float z = ps.beta * (vv[k] - ov[k]);
      float dz = 1.f, dl = 0.f;
      switch (ps.kind) {
        case DPX_PROX_NONNEG: dz = z > 0.f ? 1.f : 0.f; break;
        case DPX_PROX_L1: { const float m = fabsf(z) > le ? 1.f : 0.f; dz = m; dl = -copysignf(m, z); break; }
        case DPX_PROX_L2SQ: { const float d = 1.f / (1.f + 2.f * le); dz = d; dl = -2.f * z * d * d; break; }
        case DPX_PROX_BOX: dz = (z > ps.lo && z < ps.hi) ? 1.f : 0.f; break;
        default: break;
      }
      acc[0] += dl * gg[k];
      gg[k] *= dz;
    }
    storev<VEC>(gv + base, gg);
  }
  if (glam) {
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) atomicAdd(glam + (size_t)b * (lam_stride ? 1 : 0), acc[0] * ps.beta * ps.alpha);
  }
}

// ------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_cg_dot(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ dots, size_t per_sample) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  float acc[1] = {0.f};
  const size_t stride = (size_t)gridDim.x * blockDim.x * VEC;
  for (size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; e < per_sample; e += stride) {
    float xv[VEC], yv[VEC];
    loadv<VEC>(xv, x + (size_t)b * per_sample + e);
    loadv<VEC>(yv, y + (size_t)b * per_sample + e);
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[0] += xv[k] * yv[k];
  }
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) atomicAdd(dots + b, acc[0]);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_cg_update(float* __restrict__ x, float* __restrict__ r, const float* __restrict__ p, const float* __restrict__ q,
                const float* __restrict__ gamma, const float* __restrict__ pq, float* __restrict__ gamma_new,
                size_t per_sample) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  // a gated (converged) solve has pq = +inf (k_cg_gate): the step is skipped outright so that x and r stay frozen even
  // if the search direction has degenerated (0 * NaN would not be 0)
  const bool frozen = isinf(pq[b]);
  const float alpha = frozen ? 0.f : gamma[b] / pq[b];
  float acc[1] = {0.f};
  const size_t stride = (size_t)gridDim.x * blockDim.x * VEC;
  for (size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; e < per_sample; e += stride) {
    const size_t o = (size_t)b * per_sample + e;
    float xv[VEC], rv[VEC], pv[VEC], qv[VEC];
    loadv<VEC>(xv, x + o); loadv<VEC>(rv, r + o); loadv<VEC>(pv, p + o); loadv<VEC>(qv, q + o);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      if (!frozen) {
        xv[k] = xv[k] + alpha * pv[k];
        rv[k] = rv[k] - alpha * qv[k];
      }
      acc[0] += rv[k] * rv[k];
    }
    if (!frozen) {
      storev<VEC>(x + o, xv);
      storev<VEC>(r + o, rv);
    }
  }
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) atomicAdd(gamma_new + b, acc[0]);
}

// Device-side stop test of (P)CG (solver_cg.py:103-107 / :225-229): *done becomes (and stays) 1 once val[b] <= tol[b]
// (strict: <) holds for every b; while it is set the step scale of the following k_cg_update is forced to zero by
// pq = +inf, so the iterate is frozen at exactly the iteration where the reference breaks -- without a host round trip.
__global__ void k_cg_gate(const float* __restrict__ val, const float* __restrict__ tol, int tol_n, int strict,
                          float* __restrict__ pq, int* __restrict__ done, int batch) {
  __shared__ int all_ok;
  if (threadIdx.x == 0) all_ok = 1;
  __syncthreads();
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    const float t = tol[tol_n > 1 ? b : 0];
    const bool ok = strict ? (val[b] < t) : (val[b] <= t);
    if (!ok) all_ok = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0 && all_ok) *done = 1;
  __syncthreads();
  if (*done)
    for (int b = threadIdx.x; b < batch; b += blockDim.x) pq[b] = __int_as_float(0x7f800000);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
    k_cg_direction(float* __restrict__ p, const float* __restrict__ r, const float* __restrict__ gn,
                   const float* __restrict__ go, size_t per_sample) {
  const int b = blockIdx.y;
  const float beta = gn[b] / go[b];
  const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= per_sample) return;
  const size_t o = (size_t)b * per_sample + e;
  float pv[VEC], rv[VEC];
  loadv<VEC>(pv, p + o); loadv<VEC>(rv, r + o);
#pragma unroll
  for (int k = 0; k < VEC; ++k) pv[k] = rv[k] + beta * pv[k];
  storev<VEC>(p + o, pv);
}

// ---- launch geometry ----------------------------------------------------------------------------
inline dim3 plane_grid(size_t per_plane_elems, int vec, int planes) {
  const size_t nvec = (per_plane_elems + vec - 1) / vec;
  return dim3((unsigned)((nvec + kThreads - 1) / kThreads), (unsigned)planes, 1);
}

inline int pick_vec(const Geom& g, std::initializer_list<const void*> ptrs) {
  if (g.W % 4 != 0) return 1;
  for (const void* p : ptrs)
    if (p && !aligned16(p)) return 1;
  return 4;
}
inline int pick_vec_psi(const Geom& g, const PsiPack& psi, bool need_u, std::initializer_list<const void*> extra) {
  int v = pick_vec(g, extra);
  for (int i = 0; i < psi.n && v == 4; ++i) {
    if (!aligned16(psi.t[i].v) || (need_u && !aligned16(psi.t[i].u)) || (psi.t[i].off && !aligned16(psi.t[i].off))) v = 1;
  }
  return v;
}

}  // namespace

#define DPX_DISPATCH_VEC(vec, ...)          \
  if ((vec) == 4) {                         \
    constexpr int VEC = 4;                  \
    __VA_ARGS__;                            \
  } else {                                  \
    constexpr int VEC = 1;                  \
    __VA_ARGS__;                            \
  }

int launch_rhs(const Geom& g, const PsiPack& psi, bool hqs, float* t, cudaStream_t s) {
  const int vec = pick_vec_psi(g, psi, !hqs, {t});
  DPX_DISPATCH_VEC(vec, k_rhs<VEC><<<plane_grid(g.plane, VEC, g.P), kThreads, 0, s>>>(g, psi, hqs, t));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_spec_solve(const Geom& g, float2* spec, const float2* fb, const float* dq, int dq_batch, const float* dpsi,
                      float wid, float eps, float inv_n, RhoRef rho, cudaStream_t s) {
  const bool two = (g.splane % 2 == 0) && aligned16(spec) && (!fb || aligned16(fb)) &&
                   (!dq || (reinterpret_cast<uintptr_t>(dq) & 7u) == 0) && (!dpsi || (reinterpret_cast<uintptr_t>(dpsi) & 7u) == 0);
  if (two) {
    k_spec_solve<2><<<plane_grid(g.splane, 2, g.P), kThreads, 0, s>>>(g, spec, fb, dq, dq_batch, dpsi, wid, eps, inv_n, rho);
  } else {
    k_spec_solve<1><<<plane_grid(g.splane, 1, g.P), kThreads, 0, s>>>(g, spec, fb, dq, dq_batch, dpsi, wid, eps, inv_n, rho);
  }
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_spec_pgd(const Geom& g, float2* spec, const float2* fb, const float* dq, int dq_batch, float inv_n, RhoRef rho,
                    cudaStream_t s) {
  k_spec_pgd<<<plane_grid(g.splane, 1, g.P), kThreads, 0, s>>>(g, spec, fb, dq, dq_batch, inv_n, rho);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_mul_otf(const Geom& g, float2* spec, const float2* otf, int otf_batch, bool conj, float inv_n, cudaStream_t s) {
  k_mul_otf<<<plane_grid(g.splane, 1, g.P), kThreads, 0, s>>>(g, spec, otf, otf_batch, conj, inv_n);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_prox_dual(const Geom& g, const PsiPack& psi, const float* x, bool hqs, bool skip_external, int it,
                     float* t_fused, RhoRef rho, float* resid, cudaStream_t s) {
  const int vec = pick_vec_psi(g, psi, !hqs, {x, t_fused});
  const bool fuse = t_fused != nullptr, res = resid != nullptr;
#define DPX_PD(F, R)                                                                                              \
  DPX_DISPATCH_VEC(vec, k_prox_dual<VEC, F, R><<<plane_grid(g.plane, VEC, g.P), kThreads, 0, s>>>(g, psi, x, hqs,  \
                                                                                                   skip_external, it, \
                                                                                                   t_fused, rho, resid))
  if (fuse && res) { DPX_PD(true, true); }
  else if (fuse) { DPX_PD(true, false); }
  else if (res) { DPX_PD(false, true); }
  else { DPX_PD(false, false); }
#undef DPX_PD
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_spatial_xupdate(const Geom& g, const PsiPack& psi, bool hqs, bool vxu, const float* ktb, const float* dq,
                           int dq_batch, const float* dpsi, float wid, float eps, bool eps_delta, RhoRef rho, float* x,
                           cudaStream_t s) {
  const int vec = pick_vec_psi(g, psi, !hqs, {ktb, dq, dpsi, x});
  DPX_DISPATCH_VEC(vec, k_spatial_x<VEC><<<plane_grid(g.plane, VEC, g.P), kThreads, 0, s>>>(g, psi, hqs, vxu, ktb, dq,
                                                                                           dq_batch, dpsi, wid, eps,
                                                                                           eps_delta, rho, x));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_spatial_xupdate_bwd(const Geom& g, const float* gin, const float* x, const float* ktb, const float* dq, int dq_batch,
                               const float* dpsi, float wid, float eps, bool eps_delta, RhoRef rho, float* g_ktb, float* g_rho,
                               int g_rho_stride, cudaStream_t s) {
  PsiPack none;
  none.n = 0;
  const int vec = pick_vec_psi(g, none, false, {gin, x, ktb, dq, dpsi, g_ktb});
  if (g_rho) DPX_CUDA(cudaMemsetAsync(g_rho, 0, sizeof(float) * (g_rho_stride ? g.B : 1), s));
  DPX_DISPATCH_VEC(vec, k_spatial_x_bwd<VEC><<<plane_grid(g.plane, VEC, g.P), kThreads, 0, s>>>(
                            g, gin, x, ktb, dq, dq_batch, dpsi, wid, eps, eps_delta, rho, g_ktb, g_rho, g_rho_stride));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_pad2d(const float* in, float* out, int planes, int hi, int wi, int ho, int wo, int top, int left, cudaStream_t s) {
  const dim3 grid((unsigned)(((size_t)ho * wo + kThreads - 1) / kThreads), planes);
  k_pad2d<<<grid, kThreads, 0, s>>>(in, out, hi, wi, ho, wo, top, left);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_augment(const float* in, float* out, int planes, int h, int w, int mode, cudaStream_t s) {
  const dim3 grid((unsigned)(((size_t)h * w + kThreads - 1) / kThreads), planes);
  k_augment<<<grid, kThreads, 0, s>>>(in, out, h, w, mode);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_vxu_prox(const Geom& g, const PsiPack& psi, const float* z, int it, float* t, cudaStream_t s) {
  const int vec = pick_vec_psi(g, psi, true, {z, t});
  DPX_DISPATCH_VEC(vec, k_vxu_prox<VEC><<<plane_grid(g.plane, VEC, g.P), kThreads, 0, s>>>(g, psi, z, it, t));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_vxu_dual(const Geom& g, const PsiPack& psi, const float* z, cudaStream_t s) {
  const int vec = pick_vec_psi(g, psi, true, {z});
  DPX_DISPATCH_VEC(vec, k_vxu_dual<VEC><<<plane_grid(g.plane, VEC, g.P), kThreads, 0, s>>>(g, psi, z));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_pgd_spatial_step(const Geom& g, const float* x, const float* ktb, const float* dq, int dq_batch, RhoRef rho,
                            float* out, cudaStream_t s) {
  const int vec = pick_vec(g, {x, ktb, dq, out});
  DPX_DISPATCH_VEC(vec, k_pgd_spatial<VEC><<<plane_grid(g.plane, VEC, g.P), kThreads, 0, s>>>(g, x, ktb, dq, dq_batch,
                                                                                             rho, out));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_init_state(const Geom& g, const PsiPack& psi, const float* x, bool with_u, cudaStream_t s) {
  if (psi.n == 0) return DPX_OK;
  const int vec = pick_vec_psi(g, psi, with_u, {x});
  DPX_DISPATCH_VEC(vec, k_init_state<VEC><<<plane_grid(g.plane, VEC, g.P), kThreads, 0, s>>>(g, psi, x, with_u));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

static inline int flat_vec(size_t per_sample, std::initializer_list<const void*> ptrs) {
  if (per_sample % 4 != 0) return 1;
  for (const void* p : ptrs)
    if (p && !aligned16(p)) return 1;
  return 4;
}

int launch_prox_apply(const ProxSpec& ps, const float* v, const float* lam, int lam_stride, int it, const float* off,
                      float* out, int batch, size_t per_sample, cudaStream_t s) {
  if (ps.kind == DPX_PROX_ISO_TV) {
    DPX_REQUIRE(per_sample % 2 == 0 && off == nullptr, "iso-TV prox needs a stacked [B,2C,H,W] tensor and no offset");
    const size_t half = per_sample / 2;
    const int v2 = flat_vec(half, {v, out});
    DPX_DISPATCH_VEC(v2, k_prox_iso<VEC><<<plane_grid(half, VEC, batch), kThreads, 0, s>>>(ps, v, lam, lam_stride, it, out, half));
    DPX_LAUNCH_CHECK();
    return DPX_OK;
  }
  const int vec = flat_vec(per_sample, {v, off, out});
  DPX_DISPATCH_VEC(vec, k_prox_apply<VEC><<<plane_grid(per_sample, VEC, batch), kThreads, 0, s>>>(ps, v, lam, lam_stride, it,
                                                                                                 off, out, per_sample));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_lincomb(float* out, const float* a, const float* x, const float* b, const float* y, const float* c,
                   const float* z, int cps, int batch, size_t per_sample, cudaStream_t s) {
  const int vec = flat_vec(per_sample, {out, x, y, z});
  DPX_DISPATCH_VEC(vec, k_lincomb<VEC><<<plane_grid(per_sample, VEC, batch), kThreads, 0, s>>>(out, a, x, b, y, c, z, cps,
                                                                                              per_sample));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_grad(const float* x, float* y, int planes, int H, int W, int axis, bool adjoint, float scale, cudaStream_t s) {
  const int linop = axis == 0 ? DPX_LINOP_GRAD_H : DPX_LINOP_GRAD_W;
  const int vec = (W % 4 == 0 && aligned16(x) && aligned16(y)) ? 4 : 1;
  DPX_DISPATCH_VEC(vec, k_grad<VEC><<<plane_grid((size_t)H * W, VEC, planes), kThreads, 0, s>>>(x, y, H, W, linop, adjoint,
                                                                                               scale));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_dual_external(const float* w, const float* v_new, float* v, float* u, size_t n, cudaStream_t s) {
  const int vec = flat_vec(n, {w, v_new, v, u});
  DPX_DISPATCH_VEC(vec, k_dual_external<VEC><<<plane_grid(n, VEC, 1), kThreads, 0, s>>>(w, v_new, v, u, n));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_axpby(float* out, float a, const float* x, float b, const float* y, size_t n, cudaStream_t s) {
  const int vec = flat_vec(n, {out, x, y});
  DPX_DISPATCH_VEC(vec, k_axpby<VEC><<<plane_grid(n, VEC, 1), kThreads, 0, s>>>(out, a, x, b, y, n));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_mul(float* out, const float* x, const float* w, int w_batch, int batch, size_t per_sample, cudaStream_t s) {
  const int vec = flat_vec(per_sample, {out, x, w});
  DPX_DISPATCH_VEC(vec, k_mul<VEC><<<plane_grid(per_sample, VEC, batch), kThreads, 0, s>>>(out, x, w, w_batch, per_sample));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_resid_reduce(const float* resid, float* out, int n, int batch, cudaStream_t s) {
  k_resid_reduce<<<(n * 4 + 127) / 128, 128, 0, s>>>(resid, out, n, batch);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

static inline dim3 reduce_grid(size_t per_sample, int vec, int batch) {
  // enough CTAs to fill 148 SMs x 8 resident blocks, but no more than the data needs
  const size_t nvec = (per_sample + vec - 1) / vec;
  size_t blocks = (nvec + kThreads - 1) / kThreads;
  const size_t cap = (size_t)(148 * 8 + batch - 1) / batch;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return dim3((unsigned)blocks, (unsigned)batch, 1);
}

int launch_absmax(const float* x, float* out, int batch, size_t per_sample, cudaStream_t s) {
  DPX_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * batch, s));
  const int vec = flat_vec(per_sample, {x});
  DPX_DISPATCH_VEC(vec, k_absmax<VEC><<<reduce_grid(per_sample, VEC, batch), kThreads, 0, s>>>(x, out, per_sample));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_cg_dot(const float* x, const float* y, float* dots, int batch, size_t per_sample, cudaStream_t s) {
  DPX_CUDA(cudaMemsetAsync(dots, 0, sizeof(float) * batch, s));
  const int vec = flat_vec(per_sample, {x, y});
  DPX_DISPATCH_VEC(vec, k_cg_dot<VEC><<<reduce_grid(per_sample, VEC, batch), kThreads, 0, s>>>(x, y, dots, per_sample));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_spec_solve_bwd(const Geom& g, float2* spec, const float2* qspec, const float2* fb, const float* dq, int dq_batch,
                          const float* dpsi, float wid, float eps, float inv_n, RhoRef rho, float* g_rho, int g_rho_stride,
                          cudaStream_t s) {
  if (g_rho) DPX_CUDA(cudaMemsetAsync(g_rho, 0, sizeof(float) * (g_rho_stride ? g.B : 1), s));
  k_spec_solve_bwd<<<plane_grid(g.splane, 1, g.P), kThreads, 0, s>>>(g, spec, qspec, fb, dq, dq_batch, dpsi, wid, eps, inv_n,
                                                                      rho, g_rho, g_rho_stride);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_prox_bwd(const ProxSpec& ps, const float* v, const float* lam, int lam_stride, const float* off, const float* g,
                    float* gv, float* glam, int batch, size_t per_sample, cudaStream_t s) {
  if (glam) DPX_CUDA(cudaMemsetAsync(glam, 0, sizeof(float) * (lam_stride ? batch : 1), s));
  const int vec = flat_vec(per_sample, {v, off, g, gv});
  DPX_DISPATCH_VEC(vec, k_prox_bwd<VEC><<<reduce_grid(per_sample, VEC, batch), kThreads, 0, s>>>(ps, v, lam, lam_stride, off, g,
                                                                                                gv, glam, per_sample));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_cg_update(float* x, float* r, const float* p, const float* q, const float* gamma, const float* pq,
                     float* gamma_new, int batch, size_t per_sample, cudaStream_t s) {
  DPX_CUDA(cudaMemsetAsync(gamma_new, 0, sizeof(float) * batch, s));
  const int vec = flat_vec(per_sample, {x, r, p, q});
  DPX_DISPATCH_VEC(vec, k_cg_update<VEC><<<reduce_grid(per_sample, VEC, batch), kThreads, 0, s>>>(x, r, p, q, gamma, pq,
                                                                                                 gamma_new, per_sample));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_cg_gate(const float* val, const float* tol, int tol_n, bool strict, float* pq, int* done, int batch, cudaStream_t s) {
  k_cg_gate<<<1, 128, 0, s>>>(val, tol, tol_n, strict ? 1 : 0, pq, done, batch);
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

int launch_cg_direction(float* p, const float* r, const float* gn, const float* go, int batch, size_t per_sample,
                        cudaStream_t s) {
  const int vec = flat_vec(per_sample, {p, r});
  DPX_DISPATCH_VEC(vec, k_cg_direction<VEC><<<plane_grid(per_sample, VEC, batch), kThreads, 0, s>>>(p, r, gn, go, per_sample));
  DPX_LAUNCH_CHECK();
  return DPX_OK;
}

}  // namespace dpx
