// dpx_fused_kernels.cuh — the fused sm_100a iteration kernels (two launches per ADMM/HQS iteration).
//
//   k_col : per (plane, group of CG=4 spectrum columns): column FFT (length H)  ->  spectral solve
//           (F(K^T b) + rho*T + eps)/(sum|OTF|^2 + rho*wid + eps)  ->  inverse column FFT, in place on S.
//           Replaces cuFFT column pass + k_spec_solve + cuFFT column pass        (sum_square.py:150-152)
//   k_row : per (plane, 4 image rows = 2 row pairs): inverse row FFT (two real rows ride one complex FFT)
//           -> x;  prox + dual update of every psi term on x;  rhs of the next x-update t = sum s_i(v_i-u_i);
//           forward row FFT of t, in place on S.                                  (admm.py:49-59, hqs.py:10-16)
//           Replaces cuFFT C2R + k_prox_dual (+k_rhs) + cuFFT R2C.
//
// S is the half spectrum of the real [H,W] planes in a column-group-major layout
//     S[((p*(G+1) + g)*H + h)*CG + c]      g = k / CG, c = k % CG for spectrum column k in [0, W/2),
// plus one extra group g = G whose column c = 0 holds the Nyquist column k = W/2 (c = 1..3 unused).
// k_col therefore reads/writes one fully contiguous H*CG*8-byte tile, and k_row (4 rows at a time) touches
// 128-byte segments.  Along H the spectrum is kept in the digit-reversed order the forward FFT leaves it
// in; the solve constants are stored pre-permuted to match (see k_pack), so no reordering pass exists.
//
// DRAM traffic per real element and iteration (ADMM, one prox term): S 4R+4W (k_col) + F(K^T b) 4R + |OTF|^2 2R*
// + S 4R+4W (k_row) + u 4R+4W  = 30 B (*shared by the batch), vs 60 B with cuFFT and 24 B algorithmic.
//
// Compiles under nvcc and, with -DDPX_EMU, under g++ for the CPU emulator tests (tests/emu).
#pragma once
#include "dpx_fft_core.cuh"
#include "dpx_types.cuh"

#ifndef DPX_EMU
#define DPX_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw[]; type* name = reinterpret_cast<type*>(name##_raw)
#endif

namespace dpx {
namespace fused {

constexpr int CG = 4;            // spectrum columns per column tile
constexpr int ROWS = 4;          // image rows per row tile (2 pairs)
constexpr int kThreads = 256;

enum RowMode { ROW_FIRST = 0, ROW_MID = 1, ROW_LAST = 2 };

struct RowParams {
  int C, H;
  float2* S;
  PsiPack psi;
  int hqs;
  int it;                        // schedule column for lam
  float* x;                      // ROW_LAST: receives x
  const float2* tw;              // twiddle records of the W-tile (fft::TwiddleLayout)
};

struct ColParams {
  int C, W;
  float2* S;
  const float2* fbp;             // packed F(K^T b)      [(P*(G+1)), H/RC, CG, RC]
  const float* dqp;              // packed sum|OTF|^2    [(Cd*(G+1)), H/RC, CG, RC]
  int dq_batch;                  // 1: shared by the batch (indexed by channel), else per plane
  float wid, eps, inv_n;
  RhoRef rho;
  const float2* tw;              // twiddle records of the H-tile
};

DPX_HD size_t s_index(int p, int g, int h, int c, int H, int G) {
  return (((size_t)p * (G + 1) + g) * H + h) * CG + c;
}

DPX_HD float4 ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }
DPX_HD void stg4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

template <class TW>
struct RowSmem {
  static constexpr int RS = TW::N + 8;                    // row-buffer stride (floats): rows land on different bank halves
  static constexpr size_t TILE_BYTES = TW::SMEM_FLOAT2 * sizeof(float2);
  static constexpr size_t BYTES = TILE_BYTES + ROWS * RS * sizeof(float);
};

// ------------------------------------------------------------------------------------------------
//  Row kernel
// ------------------------------------------------------------------------------------------------
template <class TW, int MODE>
__global__ void __launch_bounds__(kThreads) k_row(RowParams P) {
  static_assert(TW::COLS == ROWS / 2, "row tile holds one complex sequence per row pair");
  constexpr int W = TW::N, NPAIR = TW::COLS, G = W / 2 / CG, RS = RowSmem<TW>::RS;
  constexpr int RA = TW::RA, MA = TW::MA;
  DPX_DYN_SMEM(float2, sm);
  float* rowbuf = reinterpret_cast<float*>(sm + TW::SMEM_FLOAT2);       // [ROWS][RS] real rows (x, then t)
  const int tid = threadIdx.x;
  const int p = blockIdx.y;
  const int r0 = blockIdx.x * ROWS;
  const int b = p / P.C;
  const int H = P.H;
  const float2* __restrict__ twA = P.tw + fft::TwiddleLayout<TW>::A_OFF;
  const float2* __restrict__ twB = P.tw + fft::TwiddleLayout<TW>::B_OFF;

  if (MODE != ROW_FIRST) {
    // ---- 1. half spectra of the 4 rows -> Z = Xa + i Xb per pair, scattered to digit-reversed positions ----
    for (int t = tid; t < G * NPAIR; t += kThreads) {
      const int pair = t % NPAIR, g = t / NPAIR;
      const float4* src = reinterpret_cast<const float4*>(P.S + s_index(p, g, r0 + 2 * pair, 0, H, G));
      const float4 a01 = src[0], a23 = src[1], b01 = src[2], b23 = src[3];      // row a: c=0..3, row b: c=0..3
      const float2 xa[4] = {make_float2(a01.x, a01.y), make_float2(a01.z, a01.w), make_float2(a23.x, a23.y), make_float2(a23.z, a23.w)};
      const float2 xb[4] = {make_float2(b01.x, b01.y), make_float2(b01.z, b01.w), make_float2(b23.x, b23.y), make_float2(b23.z, b23.w)};
#pragma unroll
      for (int c = 0; c < CG; ++c) {
        const int k = g * CG + c;
        sm[TW::phys(TW::pos_of_freq(k), pair)] = make_float2(xa[c].x - xb[c].y, xa[c].y + xb[c].x);
        if (k > 0) sm[TW::phys(TW::pos_of_freq(W - k), pair)] = make_float2(xa[c].x + xb[c].y, xb[c].x - xa[c].y);
      }
    }
    if (tid < NPAIR) {                                                          // Nyquist column k = W/2
      const size_t si = s_index(p, G, r0 + 2 * tid, 0, H, G);
      const float2 xa = P.S[si], xb = P.S[si + CG];
      sm[TW::phys(TW::pos_of_freq(W / 2), tid)] = make_float2(xa.x - xb.y, xa.y + xb.x);
    }
    __syncthreads();
    // ---- 2. inverse row FFT; its last pass writes the real rows x straight into the row buffers ---------------
    fft::smem_pass<TW, TW::RC, TW::MB, true, false>(sm, nullptr, tid, kThreads);
    __syncthreads();
    fft::smem_pass<TW, TW::RB, TW::MA, true, true>(sm, twB, tid, kThreads);
    __syncthreads();
    for (int t = tid; t < NPAIR * MA; t += kThreads) {
      const int c = t % NPAIR, j = t / NPAIR;
      const int p0 = TW::phys(j, c);
      float2 a[RA], w[RA];
#pragma unroll
      for (int m = 0; m < RA; ++m) a[m] = sm[p0 + TW::template delta<MA>(m) * NPAIR];
      fft::load_twiddles<RA>(twA + j * RA, w);
#pragma unroll
      for (int q = 1; q < RA; ++q) a[q] = fft::cmulc(a[q], w[q]);
      fft::Dft<RA, true>::run(a);
#pragma unroll
      for (int m = 0; m < RA; ++m) {
        rowbuf[(2 * c) * RS + j + m * MA] = a[m].x;
        rowbuf[(2 * c + 1) * RS + j + m * MA] = a[m].y;
      }
    }
    __syncthreads();
  }

  // ---- 3. prox / dual / next rhs: 128-bit element-wise pass over the 4 rows -------------------------------------
  for (int t = tid; t < ROWS * (W / 4); t += kThreads) {
    const int i4 = (t % (W / 4)) * 4, r = t / (W / 4);
    const size_t e = ((size_t)p * H + r0 + r) * W + i4;
    float xv[4] = {0.f, 0.f, 0.f, 0.f}, tv[4] = {0.f, 0.f, 0.f, 0.f};
    if (MODE != ROW_FIRST) {
      const float4 x4 = *reinterpret_cast<const float4*>(rowbuf + r * RS + i4);
      xv[0] = x4.x; xv[1] = x4.y; xv[2] = x4.z; xv[3] = x4.w;
    }
    for (int i = 0; i < P.psi.n; ++i) {
      const PsiTerm& tm = P.psi.t[i];
      if (MODE == ROW_FIRST) {
        const float4 v4 = ldg4(tm.v + e);
        float d[4] = {v4.x, v4.y, v4.z, v4.w};
        if (!P.hqs) {
          const float4 u4 = ldg4(tm.u + e);
          d[0] -= u4.x; d[1] -= u4.y; d[2] -= u4.z; d[3] -= u4.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) tv[k] += tm.scale * d[k];
        continue;
      }
      float off[4] = {0.f, 0.f, 0.f, 0.f}, uo[4] = {0.f, 0.f, 0.f, 0.f}, vn[4], un[4];
      if (tm.off) { const float4 o4 = ldg4(tm.off + e); off[0] = o4.x; off[1] = o4.y; off[2] = o4.z; off[3] = o4.w; }
      if (!P.hqs) { const float4 u4 = ldg4(tm.u + e); uo[0] = u4.x; uo[1] = u4.y; uo[2] = u4.z; uo[3] = u4.w; }
      const float lam = tm.lam[(size_t)b * tm.lam_stride + P.it];
      const ProxSpec ps{tm.prox, tm.alpha, tm.beta, tm.inv_beta, tm.lo, tm.hi};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float w = tm.scale * xv[k] - off[k] + uo[k];
        vn[k] = prox_wrapped(ps, w, lam, off[k]);
        un[k] = w - vn[k];
        tv[k] += tm.scale * (P.hqs ? vn[k] : vn[k] - un[k]);
      }
      if (!P.hqs) stg4(tm.u + e, make_float4(un[0], un[1], un[2], un[3]));
      if (MODE == ROW_LAST) stg4(tm.v + e, make_float4(vn[0], vn[1], vn[2], vn[3]));
    }
    if (MODE == ROW_LAST) stg4(P.x + e, make_float4(xv[0], xv[1], xv[2], xv[3]));
    else *reinterpret_cast<float4*>(rowbuf + r * RS + i4) = make_float4(tv[0], tv[1], tv[2], tv[3]);
  }
  if (MODE == ROW_LAST) return;
  __syncthreads();

  // ---- 4. forward row FFT of t; its first pass reads the row buffers ------------------------------------------------
  for (int t = tid; t < NPAIR * MA; t += kThreads) {
    const int c = t % NPAIR, j = t / NPAIR;
    const int p0 = TW::phys(j, c);
    float2 a[RA], w[RA];
#pragma unroll
    for (int m = 0; m < RA; ++m) a[m] = make_float2(rowbuf[(2 * c) * RS + j + m * MA], rowbuf[(2 * c + 1) * RS + j + m * MA]);
    fft::Dft<RA, false>::run(a);
    fft::load_twiddles<RA>(twA + j * RA, w);
#pragma unroll
    for (int q = 1; q < RA; ++q) a[q] = fft::cmul(a[q], w[q]);
#pragma unroll
    for (int m = 0; m < RA; ++m) sm[p0 + TW::template delta<MA>(m) * NPAIR] = a[m];
  }
  __syncthreads();
  fft::smem_pass<TW, TW::RB, TW::MA, false, true>(sm, twB, tid, kThreads);
  __syncthreads();
  fft::smem_pass<TW, TW::RC, TW::MB, false, false>(sm, nullptr, tid, kThreads);
  __syncthreads();

  // ---- 5. split the pair spectrum back into the two half spectra and store ---------------------------------------------
  for (int t = tid; t < G * NPAIR; t += kThreads) {
    const int pair = t % NPAIR, g = t / NPAIR;
    float2 xa[4], xb[4];
#pragma unroll
    for (int c = 0; c < CG; ++c) {
      const int k = g * CG + c;
      const float2 zk = sm[TW::phys(TW::pos_of_freq(k), pair)];
      const float2 zm = sm[TW::phys(TW::pos_of_freq((W - k) % W), pair)];
      xa[c] = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
      xb[c] = make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));
    }
    float4* dst = reinterpret_cast<float4*>(P.S + s_index(p, g, r0 + 2 * pair, 0, H, G));
    dst[0] = make_float4(xa[0].x, xa[0].y, xa[1].x, xa[1].y);
    dst[1] = make_float4(xa[2].x, xa[2].y, xa[3].x, xa[3].y);
    dst[2] = make_float4(xb[0].x, xb[0].y, xb[1].x, xb[1].y);
    dst[3] = make_float4(xb[2].x, xb[2].y, xb[3].x, xb[3].y);
  }
  if (tid < NPAIR) {
    const float2 z = sm[TW::phys(TW::pos_of_freq(W / 2), tid)];     // Z[W/2] pairs with itself
    const size_t si = s_index(p, G, r0 + 2 * tid, 0, H, G);
    P.S[si] = make_float2(z.x, 0.f);
    P.S[si + CG] = make_float2(z.y, 0.f);
  }
}

// ------------------------------------------------------------------------------------------------
//  Column kernel
// ------------------------------------------------------------------------------------------------
template <class TH>
__global__ void __launch_bounds__(kThreads) k_col(ColParams P) {
  static_assert(TH::COLS == CG, "column tile holds CG columns");
  constexpr int H = TH::N, RA = TH::RA, RC = TH::RC, MA = TH::MA;
  DPX_DYN_SMEM(float2, sm);
  const int tid = threadIdx.x;
  const int g = blockIdx.x, p = blockIdx.y;
  const int G = P.W / 2 / CG;
  const int b = p / P.C;
  float2* tile = P.S + s_index(p, g, 0, 0, H, G);
  const float2* __restrict__ twA = P.tw + fft::TwiddleLayout<TH>::A_OFF;
  const float2* __restrict__ twB = P.tw + fft::TwiddleLayout<TH>::B_OFF;

  // ---- pass A of the forward FFT, fed straight from global memory ---------------------------------------------
  for (int t = tid; t < CG * MA; t += kThreads) {
    const int c = t % CG, j = t / CG;
    const int p0 = TH::phys(j, c);
    float2 a[RA], w[RA];
#pragma unroll
    for (int m = 0; m < RA; ++m) a[m] = tile[(size_t)(j + m * MA) * CG + c];
    fft::Dft<RA, false>::run(a);
    fft::load_twiddles<RA>(twA + j * RA, w);
#pragma unroll
    for (int q = 1; q < RA; ++q) a[q] = fft::cmul(a[q], w[q]);
#pragma unroll
    for (int m = 0; m < RA; ++m) sm[p0 + TH::template delta<MA>(m) * CG] = a[m];
  }
  __syncthreads();
  fft::smem_pass<TH, TH::RB, TH::MA, false, true>(sm, twB, tid, kThreads);
  __syncthreads();

  // ---- pass C, spectral solve, inverse pass C — all on a thread-private block of RC positions ------------------------
  const float rho = P.rho.p[(size_t)b * P.rho.stride + P.rho.it];
  const int pd = P.dq_batch > 1 ? p : p % P.C;
  const float den0 = rho * P.wid + P.eps;
  for (int t = tid; t < CG * (H / RC); t += kThreads) {
    const int c = t % CG, blk = t / CG;
    const int p0 = TH::phys(blk * RC, c);
    float2 a[RC];
#pragma unroll
    for (int m = 0; m < RC; ++m) a[m] = sm[p0 + TH::template delta<1>(m) * CG];
    fft::Dft<RC, false>::run(a);
    const size_t rec = (((size_t)p * (G + 1) + g) * (H / RC) + blk) * CG + c;
    const size_t recd = (((size_t)pd * (G + 1) + g) * (H / RC) + blk) * CG + c;
    const float4* fb4 = reinterpret_cast<const float4*>(P.fbp + rec * RC);
    const float4* dq4 = reinterpret_cast<const float4*>(P.dqp + recd * RC);
    float2 f[RC];
    float d[RC];
#pragma unroll
    for (int m = 0; m < RC / 2; ++m) {
      const float4 v = fb4[m];
      f[2 * m] = make_float2(v.x, v.y); f[2 * m + 1] = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int m = 0; m < RC / 4; ++m) {
      const float4 v = dq4[m];
      d[4 * m] = v.x; d[4 * m + 1] = v.y; d[4 * m + 2] = v.z; d[4 * m + 3] = v.w;
    }
#pragma unroll
    for (int m = 0; m < RC; ++m) {
      const float den = d[m] + den0;
      a[m] = make_float2((f[m].x + rho * a[m].x + P.eps) / den * P.inv_n, (f[m].y + rho * a[m].y) / den * P.inv_n);
    }
    fft::Dft<RC, true>::run(a);
#pragma unroll
    for (int m = 0; m < RC; ++m) sm[p0 + TH::template delta<1>(m) * CG] = a[m];
  }
  __syncthreads();
  fft::smem_pass<TH, TH::RB, TH::MA, true, true>(sm, twB, tid, kThreads);
  __syncthreads();

  // ---- inverse pass A, written straight to global memory ------------------------------------------------------------------
  for (int t = tid; t < CG * MA; t += kThreads) {
    const int c = t % CG, j = t / CG;
    const int p0 = TH::phys(j, c);
    float2 a[RA], w[RA];
#pragma unroll
    for (int m = 0; m < RA; ++m) a[m] = sm[p0 + TH::template delta<MA>(m) * CG];
    fft::load_twiddles<RA>(twA + j * RA, w);
#pragma unroll
    for (int q = 1; q < RA; ++q) a[q] = fft::cmulc(a[q], w[q]);
    fft::Dft<RA, true>::run(a);
#pragma unroll
    for (int m = 0; m < RA; ++m) tile[(size_t)(j + m * MA) * CG + c] = a[m];
  }
}

// ------------------------------------------------------------------------------------------------
//  Constant packing (cold path): standard R2C layout [planes, H, W/2+1] -> k_col's record layout
// ------------------------------------------------------------------------------------------------
template <class TH, typename V>
__global__ void k_pack(const V* __restrict__ src, V* __restrict__ dst, int planes, int H, int W, int G, V zero) {
  constexpr int RC = TH::RC;
  const size_t total = (size_t)planes * (G + 1) * H * CG;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int m = (int)(i % RC);
  size_t r = i / RC;
  const int c = (int)(r % CG); r /= CG;
  const int blk = (int)(r % (H / RC)); r /= (H / RC);
  const int g = (int)(r % (G + 1));
  const int p = (int)(r / (G + 1));
  const int h = TH::freq_of_pos(blk * RC + m);
  const int k = g < G ? g * CG + c : (c == 0 ? W / 2 : -1);
  dst[i] = k >= 0 ? src[((size_t)p * H + h) * (W / 2 + 1) + k] : zero;
}

}  // namespace fused
}  // namespace dpx
