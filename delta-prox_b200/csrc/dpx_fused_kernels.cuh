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
// in; the solve constants are stored pre-permuted to match (see pack kernels), so no reordering pass exists.
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
  int C, H, W, G;                // G = (W/2)/CG
  float2* S;
  PsiPack psi;
  int hqs;
  int it;                        // schedule column for lam
  float* x;                      // ROW_LAST: receives x
  const float2* tw;              // exp(-2 pi i t / W)
};

struct ColParams {
  int C, H, W, G;
  float2* S;
  const float2* fbp;             // packed F(K^T b)      [(P*(G+1)), H/RC, CG, RC]
  const float* dqp;              // packed sum|OTF|^2    [(Cd*(G+1)), H/RC, CG, RC]
  int dq_batch;                  // 1: shared by the batch (indexed by channel), else per plane
  float wid, eps, inv_n;
  RhoRef rho;
  const float2* tw;              // exp(-2 pi i t / H)
};

DPX_HD size_t s_index(int p, int g, int h, int c, int H, int G) {
  return (((size_t)p * (G + 1) + g) * H + h) * CG + c;
}

// ------------------------------------------------------------------------------------------------
//  Row kernel
// ------------------------------------------------------------------------------------------------
template <class TW, int MODE>
__global__ void __launch_bounds__(kThreads) k_row(RowParams P) {
  static_assert(TW::COLS == ROWS / 2, "row tile holds one complex sequence per row pair");
  DPX_DYN_SMEM(float2, sm);
  const int tid = threadIdx.x;
  const int p = blockIdx.y;
  const int r0 = blockIdx.x * ROWS;
  const int b = p / P.C;
  const int W = P.W, H = P.H, G = P.G;
  constexpr int NPAIR = TW::COLS;

  if (MODE != ROW_FIRST) {
    // ---- 1. half spectra of the 4 rows -> Z = Xa + i Xb per pair, at digit-reversed positions ----
    const int ntask = G * CG * NPAIR;
    for (int t = tid; t < ntask + NPAIR; t += kThreads) {
      int g, c, pair, k;
      if (t < ntask) { c = t % CG; pair = (t / CG) % NPAIR; g = t / (CG * NPAIR); k = g * CG + c; }
      else { pair = t - ntask; g = G; c = 0; k = W / 2; }                       // Nyquist column
      const size_t si = s_index(p, g, r0 + 2 * pair, c, H, G);
      const float2 xa = P.S[si], xb = P.S[si + CG];
      sm[TW::phys(TW::pos_of_freq(k), pair)] = make_float2(xa.x - xb.y, xa.y + xb.x);
      if (k > 0 && k < W / 2) sm[TW::phys(TW::pos_of_freq(W - k), pair)] = make_float2(xa.x + xb.y, xb.x - xa.y);
    }
    __syncthreads();
    // ---- 2. inverse row FFT: sm[n] = (x_rowa[n], x_rowb[n]) ------------------------------------------
    fft::tile_fft_inverse<TW>(sm, P.tw, tid, kThreads);
  }

  // ---- 3. prox / dual / next rhs, element-wise on the 4 rows --------------------------------------------
  for (int t = tid; t < W * NPAIR; t += kThreads) {
    const int n = t % W, pair = t / W;
    const size_t ea = ((size_t)p * H + r0 + 2 * pair) * W + n, eb = ea + W;
    float xa = 0.f, xb = 0.f;
    if (MODE != ROW_FIRST) {
      const float2 z = sm[TW::phys(n, pair)];
      xa = z.x; xb = z.y;
    }
    float ta = 0.f, tb = 0.f;
    for (int i = 0; i < P.psi.n; ++i) {
      const PsiTerm& tm = P.psi.t[i];
      if (MODE == ROW_FIRST) {
        float da = tm.v[ea], db = tm.v[eb];
        if (!P.hqs) { da -= tm.u[ea]; db -= tm.u[eb]; }
        ta += tm.scale * da; tb += tm.scale * db;
        continue;
      }
      const float offa = tm.off ? tm.off[ea] : 0.f, offb = tm.off ? tm.off[eb] : 0.f;
      float wa = tm.scale * xa - offa, wb = tm.scale * xb - offb;
      if (!P.hqs) { wa += tm.u[ea]; wb += tm.u[eb]; }
      const float lam = tm.lam[(size_t)b * tm.lam_stride + P.it];
      const ProxSpec ps{tm.prox, tm.alpha, tm.beta, tm.inv_beta, tm.lo, tm.hi};
      const float va = prox_wrapped(ps, wa, lam, offa), vb = prox_wrapped(ps, wb, lam, offb);
      const float ua = wa - va, ub = wb - vb;
      if (!P.hqs) { tm.u[ea] = ua; tm.u[eb] = ub; }
      if (MODE == ROW_LAST) { tm.v[ea] = va; tm.v[eb] = vb; }
      ta += tm.scale * (P.hqs ? va : va - ua);
      tb += tm.scale * (P.hqs ? vb : vb - ub);
    }
    if (MODE == ROW_LAST) { P.x[ea] = xa; P.x[eb] = xb; }
    else sm[TW::phys(n, pair)] = make_float2(ta, tb);
  }
  if (MODE == ROW_LAST) return;
  __syncthreads();

  // ---- 4. forward row FFT of t ---------------------------------------------------------------------------
  fft::tile_fft_forward<TW>(sm, P.tw, tid, kThreads);

  // ---- 5. split the pair spectrum back into the two half spectra and store -----------------------------------
  const int ntask = G * CG * NPAIR;
  for (int t = tid; t < ntask + NPAIR; t += kThreads) {
    int g, c, pair, k;
    if (t < ntask) { c = t % CG; pair = (t / CG) % NPAIR; g = t / (CG * NPAIR); k = g * CG + c; }
    else { pair = t - ntask; g = G; c = 0; k = W / 2; }
    const float2 zk = sm[TW::phys(TW::pos_of_freq(k), pair)];
    const float2 zm = sm[TW::phys(TW::pos_of_freq((W - k) % W), pair)];
    const size_t si = s_index(p, g, r0 + 2 * pair, c, H, G);
    P.S[si] = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
    P.S[si + CG] = make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));
  }
}

// ------------------------------------------------------------------------------------------------
//  Column kernel
// ------------------------------------------------------------------------------------------------
template <class TH>
__global__ void __launch_bounds__(kThreads) k_col(ColParams P) {
  static_assert(TH::COLS == CG, "column tile holds CG columns");
  DPX_DYN_SMEM(float2, sm);
  const int tid = threadIdx.x;
  const int g = blockIdx.x, p = blockIdx.y;
  const int H = P.H, G = P.G;
  const int b = p / P.C;
  float2* tile = P.S + s_index(p, g, 0, 0, H, G);
  const float2* __restrict__ tw = P.tw;
  constexpr int RA = TH::RA, RC = TH::RC, MA = TH::MA;

  // ---- pass A of the forward FFT, fed straight from global memory ---------------------------------------------
  for (int t = tid; t < CG * MA; t += kThreads) {
    const int c = t % CG, j = t / CG;
    float2 a[RA];
#pragma unroll
    for (int m = 0; m < RA; ++m) a[m] = tile[(size_t)(j + m * MA) * CG + c];
    fft::Dft<RA, false>::run(a);
#pragma unroll
    for (int q = 1; q < RA; ++q) a[q] = fft::cmul(a[q], tw[j * q]);
#pragma unroll
    for (int m = 0; m < RA; ++m) sm[TH::phys(j + m * MA, c)] = a[m];
  }
  __syncthreads();
  fft::smem_pass<TH, TH::RB, TH::MA, false, true>(sm, tw, tid, kThreads);
  __syncthreads();

  // ---- pass C, spectral solve, inverse pass C — all on a thread-private block of RC positions ------------------------
  const float rho = P.rho.p[(size_t)b * P.rho.stride + P.rho.it];
  const int pd = P.dq_batch > 1 ? p : p % P.C;
  const float den0 = rho * P.wid + P.eps;
  for (int t = tid; t < CG * (H / RC); t += kThreads) {
    const int c = t % CG, blk = t / CG;
    float2 a[RC];
#pragma unroll
    for (int m = 0; m < RC; ++m) a[m] = sm[TH::phys(blk * RC + m, c)];
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
    for (int m = 0; m < RC; ++m) sm[TH::phys(blk * RC + m, c)] = a[m];
  }
  __syncthreads();
  fft::smem_pass<TH, TH::RB, TH::MA, true, true>(sm, tw, tid, kThreads);
  __syncthreads();

  // ---- inverse pass A, written straight to global memory ------------------------------------------------------------------
  for (int t = tid; t < CG * MA; t += kThreads) {
    const int c = t % CG, j = t / CG;
    float2 a[RA];
#pragma unroll
    for (int m = 0; m < RA; ++m) a[m] = sm[TH::phys(j + m * MA, c)];
#pragma unroll
    for (int q = 1; q < RA; ++q) a[q] = fft::cmulc(a[q], tw[j * q]);
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
