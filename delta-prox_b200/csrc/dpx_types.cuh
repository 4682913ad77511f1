// dpx_types.cuh — plain parameter structs + the prox bodies, shared by all kernels.
// Compiles under nvcc and under g++ -DDPX_EMU (tests/emu).
#pragma once

#ifdef DPX_EMU
#include "../../tests/emu/cuda_emu.h"
#else
#include <cuda_runtime.h>
#ifndef DPX_HD
#define DPX_HD __device__ __forceinline__
#endif
#endif
#include <stddef.h>

#include "../../include/dprox_b200.h"

namespace dpx {

struct Geom {
  int B, C, H, W, Wc;        // Wc = W/2+1
  int P;                     // planes = B*C
  size_t plane;              // H*W
  size_t splane;             // H*Wc
};

struct PsiTerm {
  int prox, linop;
  float scale, alpha, beta, inv_beta, lo, hi;
  float* v;
  float* u;
  const float* off;          // constant inside the linop (A x - off) or nullptr
  const float* lam;          // schedule; value for (sample b, iteration it) = lam[b*lam_stride + it]
  int lam_stride;
};
struct PsiPack {
  int n;
  PsiTerm t[DPX_MAX_PSI];
};

struct RhoRef {
  const float* p;            // value for (b, it) = p[b*stride + it]
  int stride;
  int it;
};

// `_prox` bodies: proxfn/nonneg.py:10-11, proxfn/norm.py:6-27 (+ box).
DPX_HD float prox_body(int kind, float w, float lam, float lo, float hi) {
  switch (kind) {
    case DPX_PROX_NONNEG:
      return fmaxf(w, 0.f);
    case DPX_PROX_L1:
      return copysignf(fmaxf(fabsf(w) - lam, 0.f), w);
    case DPX_PROX_L2SQ:
      return w / (1.f + 2.f * lam);
    case DPX_PROX_BOX:
      return fminf(fmaxf(w, lo), hi);
    default:
      return w;
  }
}

// ProxFn.prox wrapper chain, proxfn/base.py:12-27,55-64:
//   translated(affine(scaled(_prox, alpha), beta), off)(v, lam)
//     = 1/beta * _prox(beta*(v-off), beta*beta*lam*alpha) + off
struct ProxSpec {
  int kind;
  float alpha, beta, inv_beta, lo, hi;
};
DPX_HD float prox_wrapped(const ProxSpec& s, float v, float lam, float off) {
  const float lam_eff = s.beta * s.beta * lam * s.alpha;
  return s.inv_beta * prox_body(s.kind, s.beta * (v - off), lam_eff, s.lo, s.hi) + off;
}

}  // namespace dpx
