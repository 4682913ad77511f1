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
#define DPX_DYN_SMEM(type, name) extern __shared__ __align__(128) unsigned char name##_raw[]; type* name = reinterpret_cast<type*>(name##_raw)
#endif

namespace dpx {
namespace fused {

constexpr int CG = 4;            // spectrum columns per column tile
constexpr int ROWS = 4;          // image rows per row tile (2 pairs)
constexpr int kThreads = 256;
// column tiles so long that only ONE fits an SM (3840 and 4096 points: 139 / 148 KB) run 512 threads per CTA: 16 warps per SM
// instead of 8 (tiles up to 3072 points are co-resident two to four at a time with 256 threads each)
template <class TH> struct ColThreads { static constexpr int value = (TH::SMEM_FLOAT2 * 8 * 2 > 225 * 1024) ? 512 : kThreads; };
constexpr int kPrefetchAhead = 148 * 3;   // ~ number of k_col CTAs resident on the chip
constexpr int kPrefetchAhead4 = 148 * 4;  // ~ number of k_row CTAs resident on the chip

enum RowMode { ROW_FIRST = 0, ROW_MID = 1, ROW_LAST = 2, ROW_XONLY = 3 };   // XONLY: inverse transform -> x, nothing else (staged x-update)

struct RowParams {
  int C, H;
  float2* S;
  PsiPack psi;
  int hqs;
  int it;                        // schedule column for lam
  float* x;                      // ROW_LAST: receives x
  const float2* tw;              // twiddle records of the W-tile (fft::TwiddleLayout)
  const void* smap = nullptr;    // persistent pair kernel: tensor map of S {H*8 floats, G, pairs} in global memory, or nullptr = LDGSTS staging
  int* ctr = nullptr;            // persistent pair kernel: dynamic tile counter of this launch (zeroed), or nullptr = static round robin
  int pdl = 0;                   // row kernel launched as a programmatic dependent: wait for the previous grid before any global access
  unsigned long long* trace = nullptr;   // optional phase timestamps (DPX_TRACE), else nullptr
};

struct ColParams {
  int C, W;
  int groups;                    // column groups per (pair-)plane: W/2/CG + 1 (half spectrum) or W/CG (plane pairs)
  int bmul;                      // sample index of plane index p is bmul * (p / C): 1, or 2 for plane pairs
  float eps_im;                  // eps added to the imaginary numerator: 0, or eps when the imaginary part carries a plane
  float2* S;
  const float2* fbp;             // packed F(K^T b)      [(P*(G+1)), H/RC, CG, RC]
  const float* dqp;              // packed sum|OTF|^2    [(Cd*(G+1)), H/RC, CG, RC]
  const float* dpsp;             // packed sum_i s_i^2 |F(K_i)|^2 of the non-identity psi linops (per channel), or nullptr
  int dq_batch;                  // 1: shared by the batch (indexed by channel), else per plane
  float wid, eps, inv_n;
  RhoRef rho;
  const float2* tw;              // twiddle records of the H-tile
  unsigned long long* trace = nullptr;   // optional phase timestamps (DPX_TRACE), else nullptr
  int bulk = 0;                  // staged tile through TMA bulk copies (needs 16 bytes of shared memory behind the tile)
  int pdl = 0;                   // launched as a programmatic dependent: wait for the previous grid before touching S
};

DPX_HD size_t s_index(int p, int g, int h, int c, int H, int G) {
  return (((size_t)p * (G + 1) + g) * H + h) * CG + c;
}

// streaming global loads: data touched once per kernel must not evict the twiddle records from the small L1
// that is left next to ~220 KB of shared memory
#ifdef DPX_EMU
DPX_HD float fast_div(float a, float b) { return a / b; }
DPX_HD float2 ld_stream2(const float2* p) { return *p; }
DPX_HD float4 ld_stream4(const float4* p) { return *p; }
DPX_HD void prefetch_l2(const void*) {}
DPX_HD void trace_stamp(unsigned long long*, int, int) {}
DPX_HD void griddep_launch_dependents() {}
DPX_HD void griddep_wait() {}
#else
// Programmatic dependent launch (launch attribute programmaticStreamSerialization, DPX_PDL=1): a kernel lets the next one in the
// stream begin scheduling CTAs as soon as all of its own CTAs have started, and the next one blocks at griddep_wait() -- until the
// previous grid has completed and its memory is visible -- before it touches anything an earlier kernel of the solve wrote.  What
// overlaps with the previous kernel's tail: the launch itself, CTA scheduling, barrier set-up, the L2 prefetch of solve constants.
DPX_HD void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
DPX_HD void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
DPX_HD void trace_stamp(unsigned long long* tr, int rec, int slot) {
  if (tr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tr[(size_t)rec * 16 + slot] = t;
    if (slot == 1) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); tr[(size_t)rec * 16] = sm; }
  }
}
// pull one 128-byte line into L2 ahead of the CTA that will stream it (software pipelining across CTAs)
DPX_HD void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
DPX_HD float fast_div(float a, float b) { return __fdividef(a, b); }
DPX_HD float2 ld_stream2(const float2* p) {
  float2 r;
  asm volatile("ld.global.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
DPX_HD float4 ld_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
#endif

// asynchronous 16-byte global->shared copies (LDGSTS): the next tile streams in while the current one is computed
#ifdef DPX_EMU
DPX_HD void cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
DPX_HD void cp_async_commit() {}
DPX_HD void cp_async_wait_all() {}
#else
DPX_HD void cp_async16(void* dst, const void* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
DPX_HD void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
DPX_HD void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

// ---- bulk asynchronous copies (TMA engine, non-tensor form) + transaction barriers ---------------------------------------
// cp.async.bulk moves whole 16-byte-aligned runs global<->shared without occupying registers, LSU issue slots or the L1
// data stage; completion of loads is signalled on an mbarrier (expected-bytes transaction count), stores are tracked per
// thread in bulk groups.  SASS: UBLKCP / SYNCS.  Under DPX_EMU they are plain memcpy calls between CTA barriers.
#ifdef DPX_EMU
#define DPX_MBAR_ALL_WAIT 1
typedef unsigned long long mbar_t;
DPX_HD void mbar_init(mbar_t*, int) {}
DPX_HD void mbar_fence_init() {}
DPX_HD void mbar_expect_tx(mbar_t*, unsigned) {}
DPX_HD void mbar_wait(mbar_t*, unsigned) { __syncthreads(); }
DPX_HD void bulk_load(void* dst, const void* src, unsigned bytes, mbar_t*) { memcpy(dst, src, bytes); }
DPX_HD void bulk_store(void* dst, const void* src, unsigned bytes) { memcpy(dst, src, bytes); }
DPX_HD void bulk_commit() {}
DPX_HD void bulk_wait_read_all() {}
DPX_HD void bulk_wait_all() {}
DPX_HD void fence_async_smem() {}
DPX_HD void bulk_prefetch_l2(const void*, unsigned) {}
DPX_HD void tma_load_3d(void*, const void*, mbar_t*, int, int, int) {}
#else
// one instruction pulls a whole run into L2 (TMA engine); unlike prefetch.global.L2 it is not dropped under load
DPX_HD void bulk_prefetch_l2(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
#define DPX_MBAR_ALL_WAIT 0
typedef unsigned long long mbar_t;
DPX_HD unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
DPX_HD void mbar_init(mbar_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
DPX_HD void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DPX_HD void mbar_expect_tx(mbar_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
DPX_HD void mbar_wait(mbar_t* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
DPX_HD void bulk_load(void* dst, const void* src, unsigned bytes, mbar_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
DPX_HD void bulk_store(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
DPX_HD void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DPX_HD void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
DPX_HD void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
DPX_HD void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// tiled TMA load of one box of a 3-D tensor map that lives in global memory (UTMALDG); completion on `b`
DPX_HD void tma_load_3d(void* dst, const void* map, mbar_t* b, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(b)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
#endif

template <class TW>
struct RowSmem {
  // tile (padded, pairs interleaved) + twiddle records re-laid out for conflict-free 128-bit shared loads:
  //   twA_s[(q/2)*MA + j] = (w^{j q}, w^{j (q+1)}),  twB_s[(q/2)*MB + j] likewise
  static constexpr int TWA_F4 = (TW::RA / 2) * TW::MA, TWB_F4 = (TW::RB / 2) * TW::MB;
  static constexpr size_t TILE_BYTES = TW::SMEM_FLOAT2 * sizeof(float2);
  static constexpr size_t BYTES = TILE_BYTES + (TWA_F4 + TWB_F4) * sizeof(float4);
};

// shared-memory pass with twiddles taken from the transposed shared table `tws` ([q/2][j] float4)
template <class T, int R, int L, bool INV>
DPX_HD void smem_pass_stw(float2* sm, const float4* tws, int tid, int nthreads) {
  constexpr int M = L / R;
  constexpr int NTASK = T::COLS * T::N / R;
  constexpr bool LIN = T::template linear<M>();
  for (int task = tid; task < NTASK; task += nthreads) {
    const int c = task % T::COLS;
    const int t2 = task / T::COLS;
    const int j = t2 % M;
    const int base = (t2 / M) * L + j;
    const int p0 = T::phys(base, c);
    float2 a[R], w[R];
#pragma unroll
    for (int m = 0; m < R; ++m) a[m] = sm[LIN ? p0 + T::template delta<M>(m) * T::COLS : T::phys(base + m * M, c)];
#pragma unroll
    for (int q = 0; q < R / 2; ++q) {
      const float4 v = tws[q * M + j];
      w[2 * q] = make_float2(v.x, v.y); w[2 * q + 1] = make_float2(v.z, v.w);
    }
    if (INV) {
#pragma unroll
      for (int q = 1; q < R; ++q) a[q] = fft::cmulc(a[q], w[q]);
      fft::Dft<R, true>::run(a);
    } else {
      fft::Dft<R, false>::run(a);
#pragma unroll
      for (int q = 1; q < R; ++q) a[q] = fft::cmul(a[q], w[q]);
    }
#pragma unroll
    for (int m = 0; m < R; ++m) sm[LIN ? p0 + T::template delta<M>(m) * T::COLS : T::phys(base + m * M, c)] = a[m];
  }
}

// One psi term applied to the 2*RA image elements a thread holds after the last inverse pass
// (a[m] = (x[row a][n_m], x[row b][n_m]), n_m = j + m*MA): prox + dual update (admm.py:54-57 / hqs.py:13-15) and the
// term's contribution scale*(v - u) (ADMM) / scale*z (HQS) to the next right-hand side.  Parameters are read once.
// ACCUM: add the contribution to acc[] (several terms); otherwise overwrite a[] in place (single term).
// MODE == ROW_FIRST only forms the contribution from the stored state.
template <int MODE, bool ACCUM, int RA, int MA>
DPX_HD void row_term(const PsiTerm& tm, int hqs, int b, int it, size_t ea, size_t eb, float2 (&a)[RA], float2 (&acc)[RA]) {
  const float scale = tm.scale;
  float* __restrict__ up = tm.u;
  float* __restrict__ vp = tm.v;
  const float* __restrict__ op = tm.off;
  if (MODE == ROW_FIRST) {
#pragma unroll
    for (int m = 0; m < RA; ++m) {
      float da = vp[ea + m * MA], db = vp[eb + m * MA];
      if (!hqs) { da -= up[ea + m * MA]; db -= up[eb + m * MA]; }
      if (ACCUM) { acc[m].x += scale * da; acc[m].y += scale * db; }
      else a[m] = make_float2(scale * da, scale * db);
    }
    return;
  }
  const float lam = tm.lam[(size_t)b * tm.lam_stride + it];
  const ProxSpec ps{tm.prox, tm.alpha, tm.beta, tm.inv_beta, tm.lo, tm.hi};
#pragma unroll
  for (int m = 0; m < RA; ++m) {
    const float offa = op ? op[ea + m * MA] : 0.f, offb = op ? op[eb + m * MA] : 0.f;
    float wa = scale * a[m].x - offa, wb = scale * a[m].y - offb;
    if (!hqs) { wa += up[ea + m * MA]; wb += up[eb + m * MA]; }
    const float va = prox_wrapped(ps, wa, lam, offa), vb = prox_wrapped(ps, wb, lam, offb);
    const float ua = wa - va, ub = wb - vb;
    if (!hqs) { up[ea + m * MA] = ua; up[eb + m * MA] = ub; }
    if (MODE == ROW_LAST) { vp[ea + m * MA] = va; vp[eb + m * MA] = vb; }
    const float ca = scale * (hqs ? va : va - ua), cb = scale * (hqs ? vb : vb - ub);
    if (ACCUM) { acc[m].x += ca; acc[m].y += cb; }
    else a[m] = make_float2(ca, cb);
    // keep at most half of the dual loads in flight per thread: more would spill registers at 3 CTAs/SM
    if (m == RA / 2 - 1) asm volatile("" ::: "memory");
  }
}

// ------------------------------------------------------------------------------------------------
//  Row kernel
// ------------------------------------------------------------------------------------------------
// SINGLE: exactly one psi term (the common case) — its update runs in place on the thread's registers.
template <class TW, int MODE, bool SINGLE>
__global__ void __launch_bounds__(kThreads, (TW::N <= 2048 && SINGLE) ? 3 : 2) k_row(RowParams P) {
  static_assert(TW::COLS == ROWS / 2, "row tile holds one complex sequence per row pair");
  constexpr int W = TW::N, NPAIR = TW::COLS, G = W / 2 / CG;
  constexpr int RA = TW::RA, RB = TW::RB, MA = TW::MA, MB = TW::MB;
  DPX_DYN_SMEM(float2, sm);
  float4* twA_s = reinterpret_cast<float4*>(sm + TW::SMEM_FLOAT2);
  float4* twB_s = twA_s + RowSmem<TW>::TWA_F4;
  const int tid = threadIdx.x;
  const int p = blockIdx.y;
  const int r0 = blockIdx.x * ROWS;
  const int b = p / P.C;
  const int H = P.H;
  if (P.pdl) { griddep_launch_dependents(); griddep_wait(); }     // programmatic dependent launch: see griddep_wait

  // ---- 0. twiddle records ([q/2][j], fft::load_twiddles) -> shared memory ------------------------------------------
  {
    const float4* gA = reinterpret_cast<const float4*>(P.tw + fft::TwiddleLayout<TW>::A_OFF);
    for (int t = tid; t < RowSmem<TW>::TWA_F4; t += kThreads) twA_s[t] = gA[t];
    const float4* gB = reinterpret_cast<const float4*>(P.tw + fft::TwiddleLayout<TW>::B_OFF);
    for (int t = tid; t < RowSmem<TW>::TWB_F4; t += kThreads) twB_s[t] = gB[t];
  }

  {  // pull this CTA's own dual rows into L2 now; they are consumed two FFT passes later (step 3)
    const size_t e0 = ((size_t)p * H + r0) * W;
    for (int i = 0; i < (SINGLE ? 1 : P.psi.n); ++i) {
      const float* base = MODE == ROW_XONLY ? nullptr : (P.hqs ? (MODE == ROW_FIRST ? P.psi.t[i].v : nullptr) : P.psi.t[i].u);
      if (base) for (int o = tid * 32; o < ROWS * W; o += kThreads * 32) prefetch_l2(base + e0 + o);
    }
  }

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
    // ---- 2. inverse row FFT, passes C and B ------------------------------------------------------------------
    fft::smem_pass<TW, TW::RC, TW::MB, true, false>(sm, nullptr, tid, kThreads);
    __syncthreads();
    smem_pass_stw<TW, RB, MA, true>(sm, twB_s, tid, kThreads);
  }
  __syncthreads();

  // ---- 3. fused in registers: last inverse pass -> x;  prox / dual / next rhs on x;  first forward pass ----------
  for (int t = tid; t < NPAIR * MA; t += kThreads) {
    const int c = t % NPAIR, j = t / NPAIR;
    const int p0 = TW::phys(j, c);
    float2 a[RA];
    if (MODE != ROW_FIRST) {
#pragma unroll
      for (int m = 0; m < RA; ++m) a[m] = sm[p0 + TW::template delta<MA>(m) * NPAIR];
#pragma unroll
      for (int q = 0; q < RA / 2; ++q) {
        const float4 v = twA_s[q * MA + j];
        if (q > 0) a[2 * q] = fft::cmulc(a[2 * q], make_float2(v.x, v.y));
        a[2 * q + 1] = fft::cmulc(a[2 * q + 1], make_float2(v.z, v.w));
      }
      fft::Dft<RA, true>::run(a);                       // a[m] = (x[row a][j + m MA], x[row b][j + m MA])
    }
    const size_t ea = ((size_t)p * H + r0 + 2 * c) * W + j, eb = ea + W;
    if (MODE == ROW_LAST || MODE == ROW_XONLY) {
#pragma unroll
      for (int m = 0; m < RA; ++m) { P.x[ea + m * MA] = a[m].x; P.x[eb + m * MA] = a[m].y; }
    }
    if (MODE == ROW_XONLY) continue;
    if (SINGLE) {
      row_term<MODE, false, RA, MA>(P.psi.t[0], P.hqs, b, P.it, ea, eb, a, a);
    } else {
      float2 acc[RA];
#pragma unroll
      for (int m = 0; m < RA; ++m) acc[m] = make_float2(0.f, 0.f);
      for (int i = 0; i < P.psi.n; ++i) row_term<MODE, true, RA, MA>(P.psi.t[i], P.hqs, b, P.it, ea, eb, a, acc);
#pragma unroll
      for (int m = 0; m < RA; ++m) a[m] = acc[m];
    }
    if (MODE == ROW_LAST) continue;
    fft::Dft<RA, false>::run(a);
#pragma unroll
    for (int q = 0; q < RA / 2; ++q) {                  // twiddles re-read from shared memory: cheaper than 32 live registers
      const float4 v = twA_s[q * MA + j];
      if (q > 0) a[2 * q] = fft::cmul(a[2 * q], make_float2(v.x, v.y));
      a[2 * q + 1] = fft::cmul(a[2 * q + 1], make_float2(v.z, v.w));
    }
#pragma unroll
    for (int m = 0; m < RA; ++m) sm[p0 + TW::template delta<MA>(m) * NPAIR] = a[m];
  }
  if (MODE == ROW_LAST || MODE == ROW_XONLY) return;
  __syncthreads();

  // ---- 4. forward row FFT, passes B and C ---------------------------------------------------------------------
  smem_pass_stw<TW, RB, MA, false>(sm, twB_s, tid, kThreads);
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
//  Persistent middle row kernel (one psi term): the same arithmetic as k_row<TW, ROW_MID, true>, but every CTA
//  loops over row tiles and the NEXT tile's inputs (its 257 S segments and its 4 dual rows) are staged into
//  shared memory with cp.async while the current tile is transformed, so DRAM latency is off the critical path.
//  2 CTAs/SM:  padded tile 36 KB + S stage 32 KB + u stage 32 KB each; twiddle records come through L1.
// ------------------------------------------------------------------------------------------------
template <class TW>
struct RowPersistSmem {
  static constexpr int G = TW::N / 2 / CG;
  static constexpr int SEG_F2 = ROWS * CG;                       // float2 per S segment (4 rows x 4 columns = 128 B)
  static constexpr int STS_F2 = (G + 1) * SEG_F2;
  static constexpr int RSU = TW::N + 8;                          // staged dual-row stride (floats)
  static constexpr size_t BYTES = (TW::SMEM_FLOAT2 + STS_F2) * sizeof(float2) + ROWS * RSU * sizeof(float);
};

template <class TW>
DPX_HD void row_stage_S(const RowParams& P, int tile, float2* stS, int tid) {
  constexpr int G = TW::N / 2 / CG;
  const int tiles_per_plane = P.H / ROWS;
  const int p = tile / tiles_per_plane, r0 = (tile % tiles_per_plane) * ROWS;
  // segment g = 128 contiguous bytes (rows r0..r0+3 of group g); 8 x 16-byte chunks each
  for (int t = tid; t < (G + 1) * 8; t += kThreads) {
    const int g = t >> 3, ch = t & 7;
    cp_async16(reinterpret_cast<char*>(stS + g * (ROWS * CG)) + ch * 16,
               reinterpret_cast<const char*>(P.S + s_index(p, g, r0, 0, P.H, G)) + ch * 16);
  }
}
template <class TW>
DPX_HD void row_stage_u(const RowParams& P, int tile, float* stU, int tid) {
  constexpr int W = TW::N, RSU = RowPersistSmem<TW>::RSU;
  const int tiles_per_plane = P.H / ROWS;
  const int p = tile / tiles_per_plane, r0 = (tile % tiles_per_plane) * ROWS;
  const float* u = P.psi.t[0].u + ((size_t)p * P.H + r0) * W;
  for (int t = tid; t < ROWS * (W / 4); t += kThreads) {
    const int r = t / (W / 4), i4 = (t % (W / 4)) * 4;
    cp_async16(stU + r * RSU + i4, u + (size_t)r * W + i4);
  }
}

// prox + dual + next-rhs for the 2*RA elements a thread holds, bare `_prox` body of kind KIND (see `simple` below)
template <int KIND, int RA, int MA, bool LAST = false>
DPX_HD void mid_simple(float2 (&a)[RA], const float* ua_s, const float* ub_s, float* __restrict__ up, size_t ea, size_t eb,
                       int hqs, float lam_eff, float lo, float hi, float* __restrict__ xp = nullptr, float* __restrict__ vp = nullptr) {
#pragma unroll
  for (int m = 0; m < RA; ++m) {
    float wa = a[m].x, wb = a[m].y;
    if (LAST) { xp[ea + m * MA] = wa; xp[eb + m * MA] = wb; }          // last iteration: x leaves the kernel
    if (!hqs) { wa += ua_s[m * MA]; wb += ub_s[m * MA]; }
    const float va = prox_body(KIND, wa, lam_eff, lo, hi), vb = prox_body(KIND, wb, lam_eff, lo, hi);
    const float ua = wa - va, ub = wb - vb;
    if (!hqs) { up[ea + m * MA] = ua; up[eb + m * MA] = ub; }
    if (LAST) { vp[ea + m * MA] = va; vp[eb + m * MA] = vb; }
    else a[m] = make_float2(hqs ? va : va - ua, hqs ? vb : vb - ub);
  }
}

template <class TW>
__global__ void __launch_bounds__(kThreads, 2) k_row_mid_persist(RowParams P, int n_tiles) {
  constexpr int W = TW::N, NPAIR = TW::COLS, G = W / 2 / CG;
  constexpr int RA = TW::RA, MA = TW::MA, RSU = RowPersistSmem<TW>::RSU, SEG = ROWS * CG;
  DPX_DYN_SMEM(float2, sm);
  float2* stS = sm + TW::SMEM_FLOAT2;
  float* stU = reinterpret_cast<float*>(stS + RowPersistSmem<TW>::STS_F2);
  const int tid = threadIdx.x;
  const int H = P.H;
  const float2* __restrict__ twA = P.tw + fft::TwiddleLayout<TW>::A_OFF;
  const float2* __restrict__ twB = P.tw + fft::TwiddleLayout<TW>::B_OFF;
  const PsiTerm& tm = P.psi.t[0];
  const int hqs = P.hqs;
  if (P.pdl) { griddep_launch_dependents(); griddep_wait(); }     // programmatic dependent launch: see griddep_wait

  int tile = blockIdx.x;
  if (tile < n_tiles) {
    row_stage_S<TW>(P, tile, stS, tid);
    if (!hqs) row_stage_u<TW>(P, tile, stU, tid);
  }
  cp_async_commit();

  for (; tile < n_tiles; tile += gridDim.x) {
    const int tiles_per_plane = H / ROWS;
    const int p = tile / tiles_per_plane, r0 = (tile % tiles_per_plane) * ROWS;
    const int b = p / P.C;
    const int next = tile + gridDim.x;
    cp_async_wait_all();
    __syncthreads();                                   // staged inputs of `tile` are visible; tile buffer is free

    // ---- 1. staged half spectra -> Z = Xa + i Xb per pair, scattered to digit-reversed positions ----------------
    for (int t = tid; t < G * NPAIR * 2; t += kThreads) {
      const int half = t & 1, pair = (t >> 1) % NPAIR, g = t / (2 * NPAIR);
      const float4 av = *reinterpret_cast<const float4*>(stS + g * SEG + (2 * pair) * CG + 2 * half);       // row a, c = 2h, 2h+1
      const float4 bv = *reinterpret_cast<const float4*>(stS + g * SEG + (2 * pair + 1) * CG + 2 * half);   // row b
      const float2 xa[2] = {make_float2(av.x, av.y), make_float2(av.z, av.w)};
      const float2 xb[2] = {make_float2(bv.x, bv.y), make_float2(bv.z, bv.w)};
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int k = g * CG + 2 * half + cc;
        sm[TW::phys(TW::pos_of_freq(k), pair)] = make_float2(xa[cc].x - xb[cc].y, xa[cc].y + xb[cc].x);
        if (k > 0) sm[TW::phys(TW::pos_of_freq(W - k), pair)] = make_float2(xa[cc].x + xb[cc].y, xb[cc].x - xa[cc].y);
      }
    }
    if (tid < NPAIR) {                                                          // Nyquist column k = W/2
      const float2 xa = stS[G * SEG + (2 * tid) * CG], xb = stS[G * SEG + (2 * tid + 1) * CG];
      sm[TW::phys(TW::pos_of_freq(W / 2), tid)] = make_float2(xa.x - xb.y, xa.y + xb.x);
    }
    __syncthreads();                                   // stS consumed
    if (next < n_tiles) row_stage_S<TW>(P, next, stS, tid);
    cp_async_commit();

    // ---- 2. inverse row FFT, passes C and B ------------------------------------------------------------------
    fft::smem_pass<TW, TW::RC, TW::MB, true, false>(sm, nullptr, tid, kThreads);
    __syncthreads();
    fft::smem_pass<TW, TW::RB, TW::MA, true, true>(sm, twB, tid, kThreads);
    __syncthreads();

    // ---- 3. last inverse pass -> x;  prox / dual / next rhs in registers (dual rows from the stage);  first forward pass
    {
      const float scale = tm.scale;
      const float lam = tm.lam[(size_t)b * tm.lam_stride + P.it];
      const ProxSpec ps{tm.prox, tm.alpha, tm.beta, tm.inv_beta, tm.lo, tm.hi};
      float* __restrict__ up = tm.u;
      const float* __restrict__ op = tm.off;
      const bool simple = scale == 1.f && tm.beta == 1.f && op == nullptr &&
                          (ps.kind == DPX_PROX_NONNEG || ps.kind == DPX_PROX_L1 || ps.kind == DPX_PROX_L2SQ || ps.kind == DPX_PROX_BOX);
      const float lam_eff = lam * tm.alpha;            // beta*beta*lam*alpha with beta = 1 (bit-identical: x*1 = x)
      for (int t = tid; t < NPAIR * MA; t += kThreads) {
        const int c = t % NPAIR, j = t / NPAIR;
        const int p0 = TW::phys(j, c);
        float2 a[RA], w[RA];
#pragma unroll
        for (int m = 0; m < RA; ++m) a[m] = sm[p0 + TW::template delta<MA>(m) * NPAIR];
        fft::load_twiddles<RA, MA>(twA, j, w);
#pragma unroll
        for (int q = 1; q < RA; ++q) a[q] = fft::cmulc(a[q], w[q]);
        fft::Dft<RA, true>::run(a);
        const size_t ea = ((size_t)p * H + r0 + 2 * c) * W + j, eb = ea + W;
        const float* ua_s = stU + (2 * c) * RSU + j;
        const float* ub_s = ua_s + RSU;
        if (simple) {
          // common case (scale = beta = 1, no offset): the ProxFn wrapper chain collapses to the bare `_prox` body and the
          // switch on the prox kind is hoisted out of the element loop (it was 14 % of the kernel's instructions)
          switch (ps.kind) {
            case DPX_PROX_NONNEG: mid_simple<DPX_PROX_NONNEG, RA, MA>(a, ua_s, ub_s, up, ea, eb, hqs, lam_eff, ps.lo, ps.hi); break;
            case DPX_PROX_L1: mid_simple<DPX_PROX_L1, RA, MA>(a, ua_s, ub_s, up, ea, eb, hqs, lam_eff, ps.lo, ps.hi); break;
            case DPX_PROX_L2SQ: mid_simple<DPX_PROX_L2SQ, RA, MA>(a, ua_s, ub_s, up, ea, eb, hqs, lam_eff, ps.lo, ps.hi); break;
            default: mid_simple<DPX_PROX_BOX, RA, MA>(a, ua_s, ub_s, up, ea, eb, hqs, lam_eff, ps.lo, ps.hi); break;
          }
        } else {
#pragma unroll
          for (int m = 0; m < RA; ++m) {
            const float offa = op ? op[ea + m * MA] : 0.f, offb = op ? op[eb + m * MA] : 0.f;
            float wa = scale * a[m].x - offa, wb = scale * a[m].y - offb;
            if (!hqs) { wa += ua_s[m * MA]; wb += ub_s[m * MA]; }
            const float va = prox_wrapped(ps, wa, lam, offa), vb = prox_wrapped(ps, wb, lam, offb);
            const float ua = wa - va, ub = wb - vb;
            if (!hqs) { up[ea + m * MA] = ua; up[eb + m * MA] = ub; }
            a[m] = make_float2(scale * (hqs ? va : va - ua), scale * (hqs ? vb : vb - ub));
          }
        }
        fft::Dft<RA, false>::run(a);
#pragma unroll
        for (int q = 1; q < RA; ++q) a[q] = fft::cmul(a[q], w[q]);
#pragma unroll
        for (int m = 0; m < RA; ++m) sm[p0 + TW::template delta<MA>(m) * NPAIR] = a[m];
      }
    }
    __syncthreads();                                   // stU consumed, tile holds pass-A output
    if (next < n_tiles && !hqs) row_stage_u<TW>(P, next, stU, tid);
    cp_async_commit();

    // ---- 4. forward row FFT, passes B and C ---------------------------------------------------------------------
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
      const float2 z = sm[TW::phys(TW::pos_of_freq(W / 2), tid)];
      const size_t si = s_index(p, G, r0 + 2 * tid, 0, H, G);
      P.S[si] = make_float2(z.x, 0.f);
      P.S[si + CG] = make_float2(z.y, 0.f);
    }
  }
  cp_async_wait_all();
}

// ------------------------------------------------------------------------------------------------
//  Column kernel
// ------------------------------------------------------------------------------------------------
// STAGE: the tile is brought in by cp.async straight to its padded shared-memory positions -- the whole 64 KB tile is in
// flight at once without holding registers or L1 lines -- and pass A runs in place (measured +2.8 % on the headline workload
// against register-fed 2 x 16 loads per thread, +6.5 % at 4096 points; -1.4 ... -4 % for tiles of 1024 points and fewer, whose
// CTAs are short enough for their co-resident siblings to cover the load; profiles/README.md round 2) -- hence by tile size.
// Residency: tiles up to 1280 points leave room for FOUR co-resident CTAs (32 warps per SM) if the kernel stays within 64 registers,
// which it does without spilling: +2-4 % at 512 ... 1024 points (0.602 -> 0.615 at 1024^2); longer tiles keep three.
template <class TH, bool STAGE = (TH::N >= 2048)>
__global__ void __launch_bounds__(ColThreads<TH>::value, (TH::SMEM_FLOAT2 * sizeof(float2) * 4 <= 200 * 1024) ? 4 : ((TH::SMEM_FLOAT2 * sizeof(float2) * 3 <= 225 * 1024) ? 3 : 1)) k_col(ColParams P) {
  static_assert(TH::COLS == CG, "column tile holds CG columns");
  constexpr int H = TH::N, RA = TH::RA, RC = TH::RC, MA = TH::MA, NTH = ColThreads<TH>::value;
  DPX_DYN_SMEM(float2, sm);
  const int tid = threadIdx.x;
  // grid = (B, G+1, C): the problems of the batch are adjacent in launch order, so the sum|OTF|^2 records of
  // tile (g, channel) -- shared by the batch -- are fetched from DRAM once and hit in L2 for the other problems
  const int g = blockIdx.y;
  const int p = blockIdx.x * P.C + blockIdx.z;
  const int b = blockIdx.x * P.bmul;
  const int NG = P.groups;
  const int rec = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (tid == 0) trace_stamp(P.trace, rec, 1);
  float2* tile = P.S + ((size_t)p * NG + g) * H * CG;
  const float2* __restrict__ twA = P.tw + fft::TwiddleLayout<TH>::A_OFF;
  const float2* __restrict__ twB = P.tw + fft::TwiddleLayout<TH>::B_OFF;
  mbar_t* stage_bar = reinterpret_cast<mbar_t*>(sm + TH::SMEM_FLOAT2);    // (P.bulk) one transaction barrier behind the tile
  if (P.pdl) {
    griddep_launch_dependents();
    {  // F(K^T b) is a constant of the solve: its records can be pulled into L2 while the previous kernel drains
      const char* nf = reinterpret_cast<const char*>(P.fbp + ((size_t)p * NG + g) * H * CG);
      for (int o = tid * 128; o < H * CG * 8; o += NTH * 128) prefetch_l2(nf + o);
    }
    griddep_wait();
  }
  if (STAGE && P.bulk) {
    // TMA bulk copies: 8 rows x CG columns = 256 contiguous bytes in global memory AND at the padded position (one spare point
    // follows every 8 points), so the tile arrives through the async proxy -- no LSU wavefronts, no per-thread address stream
    if (tid == 0) { mbar_init(stage_bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) mbar_expect_tx(stage_bar, H * CG * sizeof(float2));
    __syncthreads();
    for (int i = tid; i < H / 8; i += NTH)
      bulk_load(sm + TH::pn(8 * i) * CG, tile + (size_t)i * 8 * CG, 8 * CG * sizeof(float2), stage_bar);
  } else if (STAGE) {
    // 16-byte pieces = columns (0,1) / (2,3) of one row, contiguous both in global memory and at the padded position
    for (int i = tid; i < H * CG / 2; i += NTH) {
      const int n = i >> 1, c2 = (i & 1) * 2;
      cp_async16(sm + TH::phys(n, c2), tile + (size_t)n * CG + c2);
    }
    cp_async_commit();
  }
  if (!P.pdl) {  // pull this CTA's F(K^T b) records into L2 now; they are consumed two passes later
    const char* nf = reinterpret_cast<const char*>(P.fbp + ((size_t)p * NG + g) * H * CG);
    for (int o = tid * 128; o < H * CG * 8; o += NTH * 128) prefetch_l2(nf + o);
  }

  // ---- pass A of the forward FFT: in place on the staged tile, or fed straight from global memory ---------------------
  if (STAGE) {
    if (P.bulk) {
      if (DPX_MBAR_ALL_WAIT || tid == 0) mbar_wait(stage_bar, 0);       // one poller; the barrier below publishes the tile
    } else {
      cp_async_wait_all();
    }
    __syncthreads();
    if (tid == 0) trace_stamp(P.trace, rec, 2);
    fft::smem_pass<TH, RA, H, false, true>(sm, twA, tid, NTH);
  } else {
    for (int t = tid; t < CG * MA; t += NTH) {
      const int c = t % CG, j = t / CG;
      const int p0 = TH::phys(j, c);
      float2 a[RA], w[RA];
#pragma unroll
      for (int m = 0; m < RA; ++m) a[m] = ld_stream2(tile + (size_t)(j + m * MA) * CG + c);   // streamed: the small L1 next to
      fft::Dft<RA, false>::run(a);                                                              // 217 KB of tiles is kept for twiddles
      fft::load_twiddles<RA, MA>(twA, j, w);
#pragma unroll
      for (int q = 1; q < RA; ++q) a[q] = fft::cmul(a[q], w[q]);
#pragma unroll
      for (int m = 0; m < RA; ++m) sm[p0 + TH::template delta<MA>(m) * CG] = a[m];
    }
  }
  __syncthreads();
  if (tid == 0) trace_stamp(P.trace, rec, 3);
  fft::smem_pass<TH, TH::RB, TH::MA, false, true>(sm, twB, tid, NTH);
  __syncthreads();
  if (tid == 0) trace_stamp(P.trace, rec, 4);

  // ---- pass C, spectral solve, inverse pass C — all on a thread-private block of RC positions ------------------------
  const float rho = P.rho.p[(size_t)b * P.rho.stride + P.rho.it];
  const int pd = P.dq_batch > 1 ? p : p % P.C;
  const float den0 = rho * P.wid + P.eps;
  for (int t = tid; t < CG * (H / RC); t += NTH) {
    const int c = t % CG, blk = t / CG;
    const int p0 = TH::phys(blk * RC, c);
    // constants of this block first (their DRAM/L2 latency overlaps the forward butterfly below): record [m/2][task]
    // (float4 = two spectrum values) / [m/4][task] (four diagonals), so consecutive threads read consecutive 16-byte words
    constexpr int NT = CG * (H / RC);
    const float4* fb4 = reinterpret_cast<const float4*>(P.fbp) + ((size_t)p * NG + g) * (H * CG / 2) + t;
    const float4* dq4 = reinterpret_cast<const float4*>(P.dqp) + ((size_t)pd * NG + g) * (H * CG / 4) + t;
    float2 f[RC];
    float d[RC];
#pragma unroll
    for (int m = 0; m < RC / 2; ++m) {
      const float4 v = ld_stream4(fb4 + m * NT);
      f[2 * m] = make_float2(v.x, v.y); f[2 * m + 1] = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int m = 0; m < RC / 4; ++m) {
      const float4 v = ld_stream4(dq4 + m * NT);
      d[4 * m] = v.x; d[4 * m + 1] = v.y; d[4 * m + 2] = v.z; d[4 * m + 3] = v.w;
    }
    if (P.dpsp) {              // TV-type terms: the denominator gains rho * sum_i s_i^2 |F(K_i)|^2 (sum_square.py:145-148)
      const float4* ps4 = reinterpret_cast<const float4*>(P.dpsp) + ((size_t)(p % P.C) * NG + g) * (H * CG / 4) + t;
#pragma unroll
      for (int m = 0; m < RC / 4; ++m) {
        const float4 v = ld_stream4(ps4 + m * NT);
        d[4 * m] += rho * v.x; d[4 * m + 1] += rho * v.y; d[4 * m + 2] += rho * v.z; d[4 * m + 3] += rho * v.w;
      }
    }
    float2 a[RC];
#pragma unroll
    for (int m = 0; m < RC; ++m) a[m] = sm[p0 + TH::template delta<1>(m) * CG];
    fft::Dft<RC, false>::run(a);
#pragma unroll
    for (int m = 0; m < RC; ++m) {
      // one reciprocal (MUFU.RCP, <= 1 ulp) instead of two IEEE divisions: the divisions were 19 % of the kernel's
      // instructions (profiles/README.md, v6); 1/(H W) of the unnormalised inverse transform is folded in
      const float r = fast_div(P.inv_n, d[m] + den0);
      a[m] = make_float2((f[m].x + rho * a[m].x + P.eps) * r, (f[m].y + rho * a[m].y + P.eps_im) * r);
    }
    fft::Dft<RC, true>::run(a);
#pragma unroll
    for (int m = 0; m < RC; ++m) sm[p0 + TH::template delta<1>(m) * CG] = a[m];
  }
  __syncthreads();
  if (tid == 0) trace_stamp(P.trace, rec, 5);
  fft::smem_pass<TH, TH::RB, TH::MA, true, true>(sm, twB, tid, NTH);
  __syncthreads();
  if (tid == 0) trace_stamp(P.trace, rec, 6);

  // ---- inverse pass A, written straight to global memory ------------------------------------------------------------------
  for (int t = tid; t < CG * MA; t += NTH) {
    const int c = t % CG, j = t / CG;
    const int p0 = TH::phys(j, c);
    float2 a[RA], w[RA];
#pragma unroll
    for (int m = 0; m < RA; ++m) a[m] = sm[p0 + TH::template delta<MA>(m) * CG];
    fft::load_twiddles<RA, MA>(twA, j, w);
#pragma unroll
    for (int q = 1; q < RA; ++q) a[q] = fft::cmulc(a[q], w[q]);
    fft::Dft<RA, true>::run(a);
#pragma unroll
    for (int m = 0; m < RA; ++m) tile[(size_t)(j + m * MA) * CG + c] = a[m];
  }
  if (tid == 0) trace_stamp(P.trace, rec, 7);
}

// ------------------------------------------------------------------------------------------------
//  Persistent, TMA-pipelined column kernel.  One CTA per SM loops over column tiles with THREE tile buffers in shared
//  memory: while tile i is transformed in place, tile i+1 / i+2 stream in through cp.async.bulk (256-byte runs of 8 rows
//  x CG columns land directly at their padded positions, so the FFT passes stay bank-conflict free) and tile i-1 streams
//  out, so ~128 KB per SM are in flight at all times and no global load/store of S goes through registers or the LSU.
//  Same arithmetic as k_col (bit-identical results).  H in {1024, 2048}: three padded tiles must fit 227 KB.
// ------------------------------------------------------------------------------------------------
template <class TH>
struct ColTmaCfg {
  static constexpr int NT = TH::N >= 2048 ? 512 : 256;                      // threads per CTA
  static constexpr int NBUF = 3;
  static constexpr int NCHUNK = TH::N / 8;                                  // 8 rows x CG columns = 256 bytes per bulk copy
  static constexpr unsigned CHUNK_BYTES = 8 * CG * sizeof(float2);
  static constexpr unsigned TILE_BYTES = TH::N * CG * sizeof(float2);
  static constexpr size_t BYTES = (size_t)NBUF * TH::SMEM_FLOAT2 * sizeof(float2) + NBUF * sizeof(mbar_t);
  static_assert(NCHUNK <= NT, "one bulk copy per thread");
};

template <class TH>
__global__ void __launch_bounds__(ColTmaCfg<TH>::NT, TH::N >= 2048 ? 1 : 2) k_col_tma(ColParams P, int n_tiles, int nb) {
  using Cfg = ColTmaCfg<TH>;
  constexpr int H = TH::N, RC = TH::RC, NT = Cfg::NT, NBUF = Cfg::NBUF;
  DPX_DYN_SMEM(float2, sm);
  mbar_t* full = reinterpret_cast<mbar_t*>(sm + (size_t)NBUF * TH::SMEM_FLOAT2);
  const int tid = threadIdx.x;
  const int NG = P.groups;
  const float2* __restrict__ twA = P.tw + fft::TwiddleLayout<TH>::A_OFF;
  const float2* __restrict__ twB = P.tw + fft::TwiddleLayout<TH>::B_OFF;
  // tile t -> (sample/pair bx, group g, channel ch): problems adjacent so that the shared diagonal hits in L2
  auto plane_of = [&](int t, int& g) { const int bx = t % nb; g = (t / nb) % NG; return bx * P.C + t / (nb * NG); };
  auto tile_ptr = [&](int t) { int g; const int p = plane_of(t, g); return P.S + ((size_t)p * NG + g) * H * CG; };
  auto issue_load = [&](int t, int k) {                                     // tid < NCHUNK; expect_tx already registered
    bulk_load(sm + (size_t)k * TH::SMEM_FLOAT2 + TH::pn(8 * tid) * CG, tile_ptr(t) + (size_t)tid * 8 * CG, Cfg::CHUNK_BYTES, full + k);
  };

  if (tid == 0) {
    for (int k = 0; k < NBUF; ++k) mbar_init(full + k, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int n_my = blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (tid == 0) {
    if (n_my > 0) mbar_expect_tx(full + 0, Cfg::TILE_BYTES);
    if (n_my > 1) mbar_expect_tx(full + 1, Cfg::TILE_BYTES);
  }
  __syncthreads();
  if (tid < Cfg::NCHUNK) {
    if (n_my > 0) issue_load(blockIdx.x, 0);
    if (n_my > 1) issue_load(blockIdx.x + gridDim.x, 1);
  }

  for (int i = 0; i < n_my; ++i) {
    const int k = i % NBUF;
    const int tile = blockIdx.x + i * gridDim.x;
    float2* buf = sm + (size_t)k * TH::SMEM_FLOAT2;
    int g;
    const int p = plane_of(tile, g);
    const int b = (tile % nb) * P.bmul;
    if (i + 1 < n_my) {            // constants of the next tile -> L2 (they are read in the middle of its processing)
      int gn;
      const int pn_ = plane_of(tile + gridDim.x, gn);
      const char* nf = reinterpret_cast<const char*>(P.fbp + ((size_t)pn_ * NG + gn) * H * CG);
      for (int o = tid * 128; o < H * CG * 8; o += NT * 128) prefetch_l2(nf + o);
    }
    mbar_wait(full + k, (unsigned)((i / NBUF) & 1));

    fft::smem_pass<TH, TH::RA, TH::N, false, true>(buf, twA, tid, NT);
    if (i > 0 && tid < Cfg::NCHUNK) bulk_wait_read_all();          // tile i-1 has left its buffer (the one tile i+2 will use)
    __syncthreads();
    if (tid == 0 && i + 2 < n_my) mbar_expect_tx(full + (i + 2) % NBUF, Cfg::TILE_BYTES);
    fft::smem_pass<TH, TH::RB, TH::MA, false, true>(buf, twB, tid, NT);
    __syncthreads();
    if (i + 2 < n_my && tid < Cfg::NCHUNK) issue_load(tile + 2 * gridDim.x, (i + 2) % NBUF);

    // ---- pass C, spectral solve, inverse pass C on a thread-private block of RC positions (as k_col) -----------------
    {
      const float rho = P.rho.p[(size_t)b * P.rho.stride + P.rho.it];
      const int pd = P.dq_batch > 1 ? p : p % P.C;
      const float den0 = rho * P.wid + P.eps;
      constexpr int NTASK = CG * (H / RC);
      for (int t = tid; t < NTASK; t += NT) {
        const int c = t % CG, blk = t / CG;
        const int p0 = TH::phys(blk * RC, c);
        float2 a[RC];
#pragma unroll
        for (int m = 0; m < RC; ++m) a[m] = buf[p0 + TH::template delta<1>(m) * CG];
        fft::Dft<RC, false>::run(a);
        const float4* fb4 = reinterpret_cast<const float4*>(P.fbp) + ((size_t)p * NG + g) * (H * CG / 2) + t;
        const float4* dq4 = reinterpret_cast<const float4*>(P.dqp) + ((size_t)pd * NG + g) * (H * CG / 4) + t;
        float2 f[RC];
        float d[RC];
#pragma unroll
        for (int m = 0; m < RC / 2; ++m) {
          const float4 v = fb4[m * NTASK];
          f[2 * m] = make_float2(v.x, v.y); f[2 * m + 1] = make_float2(v.z, v.w);
        }
#pragma unroll
        for (int m = 0; m < RC / 4; ++m) {
          const float4 v = dq4[m * NTASK];
          d[4 * m] = v.x; d[4 * m + 1] = v.y; d[4 * m + 2] = v.z; d[4 * m + 3] = v.w;
        }
#pragma unroll
        for (int m = 0; m < RC; ++m) {
          const float r = fast_div(P.inv_n, d[m] + den0);
          a[m] = make_float2((f[m].x + rho * a[m].x + P.eps) * r, (f[m].y + rho * a[m].y + P.eps_im) * r);
        }
        fft::Dft<RC, true>::run(a);
#pragma unroll
        for (int m = 0; m < RC; ++m) buf[p0 + TH::template delta<1>(m) * CG] = a[m];
      }
    }
    __syncthreads();
    fft::smem_pass<TH, TH::RB, TH::MA, true, true>(buf, twB, tid, NT);
    __syncthreads();
    fft::smem_pass<TH, TH::RA, TH::N, true, true>(buf, twA, tid, NT);
    fence_async_smem();                                            // generic-proxy writes -> visible to the bulk store
    __syncthreads();
    if (tid < Cfg::NCHUNK) {
      bulk_store(tile_ptr(tile) + (size_t)tid * 8 * CG, buf + TH::pn(8 * tid) * CG, Cfg::CHUNK_BYTES);
      bulk_commit();
    }
  }
  if (tid < Cfg::NCHUNK) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------------
//  Constant packing (cold path): standard R2C layout [planes, H, W/2+1] -> k_col's record layout
// ------------------------------------------------------------------------------------------------
template <class TH, typename V>
__global__ void k_pack(const V* __restrict__ src, V* __restrict__ dst, int planes, int H, int W, int G, V zero) {
  constexpr int RC = TH::RC;
  constexpr int GRP = (int)(sizeof(float4) / sizeof(V));       // values per 16-byte word: 2 (float2) or 4 (float)
  const int NT = CG * (H / RC);
  const size_t total = (size_t)planes * (G + 1) * H * CG;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // dst[tile][m / GRP][task][m % GRP],  task = blk * CG + c
  const int lane = (int)(i % GRP);
  size_t r = i / GRP;
  const int task = (int)(r % NT); r /= NT;
  const int mg = (int)(r % (RC / GRP)); r /= (RC / GRP);
  const int g = (int)(r % (G + 1));
  const int p = (int)(r / (G + 1));
  const int m = mg * GRP + lane, c = task % CG, blk = task / CG;
  const int h = TH::freq_of_pos(blk * RC + m);
  const int k = g < G ? g * CG + c : (c == 0 ? W / 2 : -1);
  dst[i] = k >= 0 ? src[((size_t)p * H + h) * (W / 2 + 1) + k] : zero;
}

// ------------------------------------------------------------------------------------------------
//  Plane-pair engine.  Two real planes that share every coefficient of the spectral solve (same channel of two
//  problems of the batch, shared rho / lam schedules) ride ONE complex 2-D transform: z = x_A + i x_B.  The solve
//  (F(K^T b) + rho T + eps)/(D + rho wid + eps) is linear with REAL coefficients, so it acts on Z = X_A + i X_B directly
//  and the Hermitian split / merge of the real-input trick (steps 1 and 5 of k_row: digit-reversed scatter with mirrored
//  indices, ~40 % of that kernel's instructions and a third of its shared-memory traffic) disappears: the first pass of
//  the row transform is fed from global memory, the last one stores to global memory, and the spectrum is a plain
//  [W/CG groups][H][CG] array of the full W columns (in the digit-reversed order the row DIF leaves them in; constants
//  are pre-permuted to match, k_packz_*).  Same DRAM traffic as the half-spectrum engine, one third fewer instructions.
//     S[((pp*G + g)*H + h)*CG + c],  G = W/CG,  pair pp = (b/2)*C + ch  holds planes (b, ch) + i (b+1, ch), b even;
//     column s = g*CG + c holds digit-reversed position (s % (W/RC))*RC + s / (W/RC)  (coalescing, see step 1).
// ------------------------------------------------------------------------------------------------
template <class TW, int MODE, bool SINGLE>
__global__ void __launch_bounds__(kThreads, TW::N <= 2048 ? 3 : 1) k_rowz(RowParams P) {
  static_assert(TW::COLS == ROWS, "row tile of the pair engine holds one complex sequence per image row");
  constexpr int W = TW::N, NSEQ = TW::COLS, G = W / CG;
  constexpr int RA = TW::RA, RB = TW::RB, RC = TW::RC, MA = TW::MA;
  static_assert((W / RC) % CG == 0, "butterfly inputs of the global-facing pass must fall into the same column of different groups");
  DPX_DYN_SMEM(float2, sm);
  const int tid = threadIdx.x;
  const int pp = blockIdx.y, h0 = blockIdx.x * ROWS;
  const int H = P.H;
  const int bq = pp / P.C;
  const int pA = 2 * bq * P.C + (pp - bq * P.C), pB = pA + P.C;
  const int b = 2 * bq;
  const float2* __restrict__ twA = P.tw + fft::TwiddleLayout<TW>::A_OFF;
  const float2* __restrict__ twB = P.tw + fft::TwiddleLayout<TW>::B_OFF;
  constexpr int NT1 = NSEQ * (W / RC);
  if (P.pdl) { griddep_launch_dependents(); griddep_wait(); }     // programmatic dependent launch: see griddep_wait

  {  // pull this CTA's dual rows into L2 now; they are consumed two FFT passes later
    const size_t eA = ((size_t)pA * H + h0) * W, eB = ((size_t)pB * H + h0) * W;
    for (int i = 0; i < (SINGLE ? 1 : P.psi.n); ++i) {
      const float* base = MODE == ROW_XONLY ? nullptr : (P.hqs ? (MODE == ROW_FIRST ? P.psi.t[i].v : nullptr) : P.psi.t[i].u);
      if (base) for (int o = tid * 32; o < ROWS * W; o += kThreads * 32) { prefetch_l2(base + eA + o); prefetch_l2(base + eB + o); }
    }
  }

  if (MODE != ROW_FIRST) {
    // ---- 1. inverse pass C fed straight from global memory.  Column storage order: position pos of the digit-reversed
    //         spectrum lives at column s = (pos % RC) * (W/RC) + pos / RC, so the RC inputs of a butterfly are W/RC columns
    //         apart and the lanes of a warp (4 columns x 4 rows, then the next group) read whole 128-byte lines.
    for (int t = tid; t < NT1; t += kThreads) {
      const int cc = t % CG, r = (t / CG) % NSEQ, gq = t / (CG * NSEQ);
      const int blk = gq * CG + cc;
      const float2* src = P.S + (((size_t)pp * G + gq) * H + h0 + r) * CG + cc;
      float2 a[RC];
#pragma unroll
      for (int m = 0; m < RC; ++m) a[m] = ld_stream2(src + (size_t)m * (W / RC / CG) * H * CG);
      fft::Dft<RC, true>::run(a);
      const int p0 = TW::phys(blk * RC, r);
#pragma unroll
      for (int m = 0; m < RC; ++m) sm[p0 + TW::template delta<1>(m) * NSEQ] = a[m];
    }
    __syncthreads();
    fft::smem_pass<TW, RB, MA, true, true>(sm, twB, tid, kThreads);
  }
  __syncthreads();

  // ---- 2. fused in registers: last inverse pass -> (x_A, x_B);  prox / dual / next rhs;  first forward pass -------------
  {
    const PsiTerm& tm = P.psi.t[0];
    const bool simple = SINGLE && MODE == ROW_MID && tm.scale == 1.f && tm.beta == 1.f && tm.off == nullptr &&
                        (tm.prox == DPX_PROX_NONNEG || tm.prox == DPX_PROX_L1 || tm.prox == DPX_PROX_L2SQ || tm.prox == DPX_PROX_BOX);
    const float lam_eff = (MODE == ROW_FIRST || MODE == ROW_XONLY) ? 0.f : tm.lam[(size_t)b * tm.lam_stride + P.it] * tm.alpha;
    for (int t = tid; t < NSEQ * MA; t += kThreads) {
      const int c = t % NSEQ, j = t / NSEQ;
      const int p0 = TW::phys(j, c);
      float2 a[RA], w[RA];
      if (MODE != ROW_FIRST) {
#pragma unroll
        for (int m = 0; m < RA; ++m) a[m] = sm[p0 + TW::template delta<MA>(m) * NSEQ];
        fft::load_twiddles<RA, MA>(twA, j, w);
#pragma unroll
        for (int q = 1; q < RA; ++q) a[q] = fft::cmulc(a[q], w[q]);
        fft::Dft<RA, true>::run(a);                     // a[m] = (x_A[h0+c][j + m MA], x_B[h0+c][j + m MA])
      }
      const size_t ea = ((size_t)pA * H + h0 + c) * W + j, eb = ((size_t)pB * H + h0 + c) * W + j;
      if (MODE == ROW_LAST || MODE == ROW_XONLY) {
#pragma unroll
        for (int m = 0; m < RA; ++m) { P.x[ea + m * MA] = a[m].x; P.x[eb + m * MA] = a[m].y; }
      }
      if (MODE == ROW_XONLY) continue;
      if (simple) {
        float* __restrict__ up = tm.u;
        switch (tm.prox) {
          case DPX_PROX_NONNEG: mid_simple<DPX_PROX_NONNEG, RA, MA>(a, up + ea, up + eb, up, ea, eb, P.hqs, lam_eff, tm.lo, tm.hi); break;
          case DPX_PROX_L1: mid_simple<DPX_PROX_L1, RA, MA>(a, up + ea, up + eb, up, ea, eb, P.hqs, lam_eff, tm.lo, tm.hi); break;
          case DPX_PROX_L2SQ: mid_simple<DPX_PROX_L2SQ, RA, MA>(a, up + ea, up + eb, up, ea, eb, P.hqs, lam_eff, tm.lo, tm.hi); break;
          default: mid_simple<DPX_PROX_BOX, RA, MA>(a, up + ea, up + eb, up, ea, eb, P.hqs, lam_eff, tm.lo, tm.hi); break;
        }
      } else if (SINGLE) {
        row_term<MODE, false, RA, MA>(tm, P.hqs, b, P.it, ea, eb, a, a);
      } else {
        float2 acc[RA];
#pragma unroll
        for (int m = 0; m < RA; ++m) acc[m] = make_float2(0.f, 0.f);
        for (int i = 0; i < P.psi.n; ++i) row_term<MODE, true, RA, MA>(P.psi.t[i], P.hqs, b, P.it, ea, eb, a, acc);
#pragma unroll
        for (int m = 0; m < RA; ++m) a[m] = acc[m];
      }
      if (MODE == ROW_LAST) continue;
      fft::Dft<RA, false>::run(a);
#ifndef DPX_EMU
      asm volatile("" ::: "memory");                    // re-read the twiddle record (L1 hit) instead of keeping 32 registers live
#endif
      fft::load_twiddles<RA, MA>(twA, j, w);
#pragma unroll
      for (int q = 1; q < RA; ++q) a[q] = fft::cmul(a[q], w[q]);
#pragma unroll
      for (int m = 0; m < RA; ++m) sm[p0 + TW::template delta<MA>(m) * NSEQ] = a[m];
    }
  }
  if (MODE == ROW_LAST || MODE == ROW_XONLY) return;
  __syncthreads();

  // ---- 3. forward pass B in shared memory, forward pass C stored straight to global memory ----------------------------
  fft::smem_pass<TW, RB, MA, false, true>(sm, twB, tid, kThreads);
  __syncthreads();
  for (int t = tid; t < NT1; t += kThreads) {
    const int cc = t % CG, r = (t / CG) % NSEQ, gq = t / (CG * NSEQ);
    const int blk = gq * CG + cc;
    const int p0 = TW::phys(blk * RC, r);
    float2 a[RC];
#pragma unroll
    for (int m = 0; m < RC; ++m) a[m] = sm[p0 + TW::template delta<1>(m) * NSEQ];
    fft::Dft<RC, false>::run(a);
    float2* dst = P.S + (((size_t)pp * G + gq) * H + h0 + r) * CG + cc;
#pragma unroll
    for (int m = 0; m < RC; ++m) dst[(size_t)m * (W / RC / CG) * H * CG] = a[m];
  }
}

// ------------------------------------------------------------------------------------------------
//  Persistent middle row kernel of the plane-pair engine (one psi term).  Same arithmetic as k_rowz<.., ROW_MID, true>
//  on tiles of ZR = 2 image rows; every CTA loops over tiles and the NEXT tile's spectrum rows and dual rows are staged
//  into shared memory by the TMA engine (cp.async.bulk + transaction barriers: no registers, no LSU issue slots) while the
//  current tile is transformed (2 CTAs/SM: padded tile 37 KB + 32 KB + 32 KB).
// ------------------------------------------------------------------------------------------------
constexpr int ZR = 2;
template <class TW>
struct RowZPersistSmem {
  static constexpr int G = TW::N / CG;
  static constexpr int STS_F2 = G * ZR * CG;                     // staged spectrum rows: [g][r][c]
  static constexpr int RSU = TW::N + 16;                         // staged dual-row stride (floats): rows land in disjoint banks
  static constexpr int STS_OFF = (TW::SMEM_FLOAT2 + 15) / 16 * 16;   // float2 offset of the stage: 128-byte aligned (TMA tensor loads land there)
  static constexpr size_t BYTES = (STS_OFF + STS_F2) * sizeof(float2) + 2 * ZR * RSU * sizeof(float) + 2 * sizeof(mbar_t) + 16;   // + next-tile slot
  // rows up to 1024 points leave room for a third co-resident CTA (24 instead of 16 warps per SM) at 85 registers per thread
  // ... and a fourth up to 768 points (64 registers; 0.577 -> 0.613 at 768^2; at 1024 points the 64-register version spills and loses)
  // rows from 3072 points (tile + stages > 113 KB) fit ONE CTA per SM: it runs 512 threads, so the SM keeps the 16 warps that two
  // co-resident 256-thread CTAs give the 2048-point rows (8 warps per SM measured 0.47 of the roofline at 4096^2, 0.40 at 3840 x 2160)
  static constexpr bool SOLO = BYTES * 2 > 225 * 1024;
  static constexpr int THREADS = SOLO ? 512 : kThreads;
  static constexpr int CTAS_PER_SM = SOLO ? 1 : ((TW::N <= 768 && BYTES * 4 <= 225 * 1024) ? 4 : ((TW::N <= 1024 && BYTES * 3 <= 225 * 1024) ? 3 : 2));
};

struct RowZTile { int pp, h0, pA, pB; };
DPX_HD RowZTile rowz_tile(const RowParams& P, int tile, int tiles_per_pair, int n_tiles) {
  RowZTile t;
  const int seq = tile / tiles_per_pair;
  t.h0 = (tile - seq * tiles_per_pair) * ZR;
  t.pp = seq;
  (void)n_tiles;
  const int bq = t.pp / P.C;
  t.pA = 2 * bq * P.C + (t.pp - bq * P.C); t.pB = t.pA + P.C;
  return t;
}
// staging of a tile's inputs: G spectrum segments of 64 bytes (rows h0, h0+1 of one column group; cp.async) and the
// 2 x ZR dual rows of 4 W bytes (bulk copies by the TMA engine, completion counted on a transaction barrier)
template <class TW>
DPX_HD void rowz_stage_S(const RowParams& P, const RowZTile& t, float2* stS, int tid) {
  // 64-byte segments: too small for the TMA engine to be efficient (measured: 512 bulk copies per tile were slower than
  // LDGSTS), so the spectrum rows are staged with 16-byte cp.async; the 8 KB dual rows below go through the TMA engine
  constexpr int G = TW::N / CG, SEG16 = ZR * CG * 8 / 16;
  const char* base = reinterpret_cast<const char*>(P.S + ((size_t)t.pp * G * P.H + t.h0) * CG);
  for (int i = tid; i < G * SEG16; i += RowZPersistSmem<TW>::THREADS) {
    const int g = i / SEG16, ch = i % SEG16;
    cp_async16(reinterpret_cast<char*>(stS + g * (ZR * CG)) + ch * 16, base + (size_t)g * P.H * CG * sizeof(float2) + ch * 16);
  }
}
// the same rows through the TMA engine: one tensor-map box {2 rows x CG columns = 64 B, up to 256 column groups} per instruction,
// landing in the [g][r][c] order of the stage; no LSU wavefronts and no per-thread address stream (one elected thread)
constexpr int rowz_tma_box(int G) {                 // column groups per TMA box: the largest divisor of G that a box dimension can hold
  int b = G < 256 ? G : 256;
  while (G % b) --b;
  return b;
}
template <class TW>
DPX_HD void rowz_stage_S_tma(const RowParams& P, const RowZTile& t, float2* stS, mbar_t* bar) {
  constexpr int G = TW::N / CG, BOX = rowz_tma_box(G);
  mbar_expect_tx(bar, G * ZR * CG * sizeof(float2));
  for (int g0 = 0; g0 < G; g0 += BOX) tma_load_3d(stS + g0 * (ZR * CG), P.smap, bar, t.h0 * CG * 2, g0, t.pp);
}
template <class TW>
DPX_HD void rowz_stage_u(const RowParams& P, const RowZTile& t, float* stU, mbar_t* bar, int tid, const float* base = nullptr) {
  constexpr int W = TW::N, RSU = RowZPersistSmem<TW>::RSU;
  if (!base) base = P.psi.t[0].u;
  if (tid < 2 * ZR) {
    const int pl = tid / ZR, r = tid % ZR;                         // row = plane * ZR + r
    bulk_load(stU + tid * RSU, base + ((size_t)(pl ? t.pB : t.pA) * P.H + t.h0 + r) * W, W * sizeof(float), bar);
  }
}

// LAST: the final iteration of a call -- x and v are written out next to u and there is no forward transform (the same staged,
// persistent pipeline instead of the non-persistent k_rowz<ROW_LAST>: 350 -> ~200 us per two problems at 2048^2)
// PM (persist mode): PM_MID; PM_LAST; PM_XONLY = inverse transform -> x only (staged x-update: an external prox or a stencil
// prox follows); PM_FIRST = forward transform of scale * (v - u) of the single psi term (the rows of v arrive through the dual
// stage; the dual itself, read once per solve, straight from global memory).  The last two replace the non-persistent, unstaged
// k_rowz<ROW_XONLY / ROW_FIRST> on the staged x-update path (TV objectives, deep priors) and at the start of every solve.
enum PersistMode { PM_MID = 0, PM_LAST = 1, PM_XONLY = 2, PM_FIRST = 3 };
template <class TW, int PM = PM_MID>
__global__ void __launch_bounds__(RowZPersistSmem<TW>::THREADS, RowZPersistSmem<TW>::CTAS_PER_SM) k_rowz_mid_persist(RowParams P, int n_tiles) {
  constexpr bool LAST = PM == PM_LAST || PM == PM_XONLY;       // no forward transform, x leaves the kernel
  constexpr bool XONLY = PM == PM_XONLY, FIRST = PM == PM_FIRST;
  static_assert(TW::COLS == ZR, "tile holds one complex sequence per image row");
  constexpr int W = TW::N, NSEQ = ZR, G = W / CG;
  constexpr int RA = TW::RA, RB = TW::RB, RC = TW::RC, MA = TW::MA, RSU = RowZPersistSmem<TW>::RSU;
  constexpr int NT1 = NSEQ * (W / RC), NTH = RowZPersistSmem<TW>::THREADS;
  constexpr unsigned U_BYTES = 2 * ZR * W * sizeof(float);
  static_assert((W / RC) % CG == 0, "butterfly inputs of the global-facing pass fall into the same column of different groups");
  DPX_DYN_SMEM(float2, sm);
  float2* stS = sm + RowZPersistSmem<TW>::STS_OFF;
  float* stU = reinterpret_cast<float*>(stS + RowZPersistSmem<TW>::STS_F2);
  mbar_t* bars = reinterpret_cast<mbar_t*>(stU + 2 * ZR * RSU);    // [1]: dual stage
  const int tid = threadIdx.x;
  const int H = P.H, tpp = H / ZR;
  const float2* __restrict__ twA = P.tw + fft::TwiddleLayout<TW>::A_OFF;
  const float2* __restrict__ twB = P.tw + fft::TwiddleLayout<TW>::B_OFF;
  const PsiTerm& tm = P.psi.t[0];
  const int hqs = P.hqs;
  const bool ust = FIRST || (!XONLY && !hqs);       // rows staged next to the spectrum: the dual (v for PM_FIRST)
  const bool sst = !FIRST;                          // spectrum rows staged (PM_FIRST has no inverse transform)

  if (tid == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    mbar_fence_init();
  }
  if (P.pdl) { griddep_launch_dependents(); griddep_wait(); }
  __syncthreads();
  int tile = blockIdx.x;
  RowZTile cur = rowz_tile(P, tile < n_tiles ? tile : 0, tpp, n_tiles);
  if (tile < n_tiles) {
    if (tid == 0 && ust) mbar_expect_tx(bars + 1, U_BYTES);
    __syncthreads();
    if (sst) {
      if (P.smap) { if (tid == 0) rowz_stage_S_tma<TW>(P, cur, stS, bars + 0); }
      else rowz_stage_S<TW>(P, cur, stS, tid);
    }
    if (ust) rowz_stage_u<TW>(P, cur, stU, bars + 1, tid, FIRST ? tm.v : tm.u);
  }
  cp_async_commit();

  // tiles after the first come from a device counter when one is given: SMs differ by ~13 % in per-tile time (distance to
  // the L2 slices), and a static round robin leaves the fast ones idle for the last 10 % of the kernel (DPX_TRACE timeline)
  volatile int* s_next = reinterpret_cast<volatile int*>(bars + 2);
  unsigned phase = 0;
  int next = tile + gridDim.x;
  for (; tile < n_tiles; tile = next, phase ^= 1u) {
    const int pp = cur.pp, h0 = cur.h0, pA = cur.pA, pB = cur.pB;
    const int b = 2 * (pp / P.C);
    if (tid == 0) {
      trace_stamp(P.trace, tile, 1);
#ifndef DPX_EMU
      if (P.ctr) *s_next = (int)gridDim.x + atomicAdd(P.ctr, 1);
      else
#endif
        *s_next = tile + (int)gridDim.x;
    }
    cp_async_wait_all();
    if (DPX_MBAR_ALL_WAIT || tid == 0) {                                           // one poller; the barrier below publishes it
      if (P.smap && sst) mbar_wait(bars + 0, phase);
      if (ust) mbar_wait(bars + 1, phase);
    }
    __syncthreads();                                   // staged inputs are visible; the tile buffer is free
    next = *s_next;
    const RowZTile nxt = rowz_tile(P, next < n_tiles ? next : tile, tpp, n_tiles);
    if (tid == 0) trace_stamp(P.trace, tile, 2);
    if (tid == 0 && next < n_tiles && ust) mbar_expect_tx(bars + 1, U_BYTES);   // copies are issued behind later barriers

    // ---- 1. inverse pass C out of the staged spectrum rows (column storage order: see k_rowz) ----------------------------
    if (!FIRST) {
    for (int t = tid; t < NT1; t += NTH) {
      const int cc = t % CG, r = (t / CG) % NSEQ, gq = t / (CG * NSEQ);
      const int blk = gq * CG + cc;
      const float2* src = stS + gq * (ZR * CG) + r * CG + cc;
      float2 a[RC];
#pragma unroll
      for (int m = 0; m < RC; ++m) a[m] = src[m * (W / RC / CG) * (ZR * CG)];
      fft::Dft<RC, true>::run(a);
      const int p0 = TW::phys(blk * RC, r);
#pragma unroll
      for (int m = 0; m < RC; ++m) sm[p0 + TW::template delta<1>(m) * NSEQ] = a[m];
    }
    __syncthreads();                                   // stS consumed
    if (tid == 0) trace_stamp(P.trace, tile, 3);
    if (next < n_tiles) {
      if (P.smap) { if (tid == 0) rowz_stage_S_tma<TW>(P, nxt, stS, bars + 0); }
      else rowz_stage_S<TW>(P, nxt, stS, tid);
    }
    cp_async_commit();
    fft::smem_pass<TW, RB, MA, true, true>(sm, twB, tid, NTH);
    __syncthreads();
    }
    if (tid == 0) trace_stamp(P.trace, tile, 4);

    // ---- 2. last inverse pass -> (x_A, x_B);  prox / dual / next rhs in registers (dual rows from the stage);  first forward pass
    {
      const float scale = tm.scale;
      const float lam = (XONLY || FIRST) ? 0.f : tm.lam[(size_t)b * tm.lam_stride + P.it];
      const ProxSpec ps{tm.prox, tm.alpha, tm.beta, tm.inv_beta, tm.lo, tm.hi};
      float* __restrict__ up = tm.u;
      const float* __restrict__ op = tm.off;
      const bool simple = scale == 1.f && tm.beta == 1.f && op == nullptr &&
                          (ps.kind == DPX_PROX_NONNEG || ps.kind == DPX_PROX_L1 || ps.kind == DPX_PROX_L2SQ || ps.kind == DPX_PROX_BOX);
      const float lam_eff = lam * tm.alpha;
      for (int t = tid; t < NSEQ * MA; t += NTH) {
        const int c = t % NSEQ, j = t / NSEQ;
        const int p0 = TW::phys(j, c);
        float2 a[RA], w[RA];
        fft::load_twiddles<RA, MA>(twA, j, w);
        if (!FIRST) {
#pragma unroll
          for (int m = 0; m < RA; ++m) a[m] = sm[p0 + TW::template delta<MA>(m) * NSEQ];
#pragma unroll
          for (int q = 1; q < RA; ++q) a[q] = fft::cmulc(a[q], w[q]);
          fft::Dft<RA, true>::run(a);
        }
        const size_t ea = ((size_t)pA * H + h0 + c) * W + j, eb = ((size_t)pB * H + h0 + c) * W + j;
        const float* ua_s = stU + c * RSU + j;
        const float* ub_s = stU + (ZR + c) * RSU + j;
        if (XONLY) {
          float* __restrict__ xp = P.x;
#pragma unroll
          for (int m = 0; m < RA; ++m) { xp[ea + m * MA] = a[m].x; xp[eb + m * MA] = a[m].y; }
        } else if (FIRST) {
          // right-hand side of the first x-update from the stored state: scale * (v - u) (ADMM) / scale * v (HQS); v is staged
#pragma unroll
          for (int m = 0; m < RA; ++m) {
            float da = ua_s[m * MA], db = ub_s[m * MA];
            if (!hqs) { da -= up[ea + m * MA]; db -= up[eb + m * MA]; }
            a[m] = make_float2(scale * da, scale * db);
          }
        } else if (simple) {
          switch (ps.kind) {
            case DPX_PROX_NONNEG: mid_simple<DPX_PROX_NONNEG, RA, MA, LAST>(a, ua_s, ub_s, up, ea, eb, hqs, lam_eff, ps.lo, ps.hi, P.x, tm.v); break;
            case DPX_PROX_L1: mid_simple<DPX_PROX_L1, RA, MA, LAST>(a, ua_s, ub_s, up, ea, eb, hqs, lam_eff, ps.lo, ps.hi, P.x, tm.v); break;
            case DPX_PROX_L2SQ: mid_simple<DPX_PROX_L2SQ, RA, MA, LAST>(a, ua_s, ub_s, up, ea, eb, hqs, lam_eff, ps.lo, ps.hi, P.x, tm.v); break;
            default: mid_simple<DPX_PROX_BOX, RA, MA, LAST>(a, ua_s, ub_s, up, ea, eb, hqs, lam_eff, ps.lo, ps.hi, P.x, tm.v); break;
          }
        } else {
          float* __restrict__ xp = P.x;
          float* __restrict__ vp = tm.v;
#pragma unroll
          for (int m = 0; m < RA; ++m) {
            if (LAST) { xp[ea + m * MA] = a[m].x; xp[eb + m * MA] = a[m].y; }
            const float offa = op ? op[ea + m * MA] : 0.f, offb = op ? op[eb + m * MA] : 0.f;
            float wa = scale * a[m].x - offa, wb = scale * a[m].y - offb;
            if (!hqs) { wa += ua_s[m * MA]; wb += ub_s[m * MA]; }
            const float va = prox_wrapped(ps, wa, lam, offa), vb = prox_wrapped(ps, wb, lam, offb);
            const float ua = wa - va, ub = wb - vb;
            if (!hqs) { up[ea + m * MA] = ua; up[eb + m * MA] = ub; }
            if (LAST) { vp[ea + m * MA] = va; vp[eb + m * MA] = vb; }
            else a[m] = make_float2(scale * (hqs ? va : va - ua), scale * (hqs ? vb : vb - ub));
          }
        }
        if (!LAST) {
          fft::Dft<RA, false>::run(a);
#pragma unroll
          for (int q = 1; q < RA; ++q) a[q] = fft::cmul(a[q], w[q]);   // twiddles stay in registers (2 CTAs/SM leave 128 each)
#pragma unroll
          for (int m = 0; m < RA; ++m) sm[p0 + TW::template delta<MA>(m) * NSEQ] = a[m];
        }
      }
    }
    __syncthreads();                                   // stU consumed, tile holds pass-A output
    if (tid == 0) trace_stamp(P.trace, tile, 5);
    if (next < n_tiles && ust) rowz_stage_u<TW>(P, nxt, stU, bars + 1, tid, FIRST ? tm.v : tm.u);

    // ---- 3. forward pass B in shared memory, forward pass C stored straight to global memory ---------------------------
    if (!LAST) {
    fft::smem_pass<TW, RB, MA, false, true>(sm, twB, tid, NTH);
    __syncthreads();
    if (tid == 0) trace_stamp(P.trace, tile, 6);
    for (int t = tid; t < NT1; t += NTH) {
      const int cc = t % CG, r = (t / CG) % NSEQ, gq = t / (CG * NSEQ);
      const int blk = gq * CG + cc;
      const int p0 = TW::phys(blk * RC, r);
      float2 a[RC];
#pragma unroll
      for (int m = 0; m < RC; ++m) a[m] = sm[p0 + TW::template delta<1>(m) * NSEQ];
      fft::Dft<RC, false>::run(a);
      float2* dst = P.S + (((size_t)pp * G + gq) * H + h0 + r) * CG + cc;
#pragma unroll
      for (int m = 0; m < RC; ++m) dst[(size_t)m * (W / RC / CG) * H * CG] = a[m];
    }
    }
    cur = nxt;
    if (tid == 0) trace_stamp(P.trace, tile, 7);
  }
  cp_async_wait_all();
}

// Constant packing for the pair engine (cold path).  Standard R2C half spectra -> full-spectrum records of pair pp in
// k_col's per-thread record order, columns in the row transform's storage order:
//   fbz = F(K^T b)_A + i F(K^T b)_B   (bins k > W/2 by Hermitian symmetry),  dqz = sum|OTF|^2 (even symmetry).
// The tile radices are run-time arguments here: templating on (TH, TW) would instantiate one kernel per H x W combination.
struct PackGeom {
  int H, W;
  int hRA, hRB, hRC;             // radices of the column (H) transform
  int wRA, wRB, wRC;             // radices of the row (W) transform
};
DPX_HD int freq_of_pos_rt(int p, int N, int RA, int RB) {   // fft::Tile::freq_of_pos with run-time radices
  const int MA = N / RA, MB = MA / RB;
  return p / MA + RA * ((p % MA) / MB) + RA * RB * (p % MB);
}
static __global__ void k_packz_fb(const float2* __restrict__ src, float2* __restrict__ dst, int pairs, int C, PackGeom q) {
  const int H = q.H, W = q.W, RC = q.hRC;
  const int NT = CG * (H / RC), G = W / CG, Wc = W / 2 + 1;
  const size_t total = (size_t)pairs * G * H * CG;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int lane = (int)(i % 2);                     // dst[tile][m / 2][task][m % 2]
  size_t r = i / 2;
  const int task = (int)(r % NT); r /= NT;
  const int mg = (int)(r % (RC / 2)); r /= (RC / 2);
  const int g = (int)(r % G);
  const int pp = (int)(r / G);
  const int m = mg * 2 + lane, c = task % CG, blk = task / CG;
  const int sc = g * CG + c;                          // storage column -> digit-reversed position -> frequency
  const int h = freq_of_pos_rt(blk * RC + m, H, q.hRA, q.hRB);
  const int k = freq_of_pos_rt((sc % (W / q.wRC)) * q.wRC + sc / (W / q.wRC), W, q.wRA, q.wRB);
  const int bq = pp / C, pA = 2 * bq * C + (pp - bq * C), pB = pA + C;
  float2 fa, fb;
  if (k < Wc) {
    fa = src[((size_t)pA * H + h) * Wc + k];
    fb = src[((size_t)pB * H + h) * Wc + k];
  } else {
    const int hm = (H - h) % H, km = W - k;
    fa = src[((size_t)pA * H + hm) * Wc + km]; fa.y = -fa.y;
    fb = src[((size_t)pB * H + hm) * Wc + km]; fb.y = -fb.y;
  }
  dst[i] = make_float2(fa.x - fb.y, fa.y + fb.x);
}
// Same packing, organised around the SOURCE rows: a CTA stages the 8 rows one (pair, block, m-pair) needs -- rows h(m), h(m+1) and
// their mirrors H - h of both planes -- with fully coalesced loads and writes, per column group, the 64-byte run
// (c = 0..3) x (m % 2) of the record layout.  k_packz_fb reads one scattered 8-byte element per thread (every load its own
// 32-byte sector, 256 rows apart): 318 us per two problems at 2048^2, 10 % of an end-to-end step; this form is bandwidth-bound.
// grid (RC / 2, H / RC, pairs), dynamic shared memory 8 * (W / 2 + 1) * 8 bytes.
#ifndef DPX_EMU
static __global__ void k_packz_fb_rows(const float2* __restrict__ src, float2* __restrict__ dst, int pairs, int C, PackGeom q) {
  extern __shared__ __align__(16) unsigned char rows_raw[];
  float2* rows = reinterpret_cast<float2*>(rows_raw);              // [plane][mirror][m % 2][Wc]
  const int H = q.H, W = q.W, RC = q.hRC;
  const int NT = CG * (H / RC), G = W / CG, Wc = W / 2 + 1;
  const int mg = blockIdx.x, blk = blockIdx.y, pp = blockIdx.z;
  const int bq = pp / C, pA = 2 * bq * C + (pp - bq * C), pB = pA + C;
  for (int r = 0; r < 8; ++r) {
    const int pl = r >> 2, mir = (r >> 1) & 1, m2 = r & 1;
    const int h = freq_of_pos_rt(blk * RC + mg * 2 + m2, H, q.hRA, q.hRB);
    const float2* s = src + ((size_t)(pl ? pB : pA) * H + (mir ? (H - h) % H : h)) * Wc;
    for (int k = threadIdx.x; k < Wc; k += blockDim.x) rows[r * Wc + k] = s[k];
  }
  __syncthreads();
  const int WR = W / q.wRC;
  for (int idx = threadIdx.x; idx < G * 8; idx += blockDim.x) {
    const int g = idx >> 3, c = (idx >> 1) & 3, m2 = idx & 1;
    const int sc = g * CG + c;
    const int k = freq_of_pos_rt((sc % WR) * q.wRC + sc / WR, W, q.wRA, q.wRB);
    float2 fa, fb;
    if (k < Wc) {
      fa = rows[(0 + m2) * Wc + k];
      fb = rows[(4 + m2) * Wc + k];
    } else {
      fa = rows[(2 + m2) * Wc + (W - k)]; fa.y = -fa.y;
      fb = rows[(6 + m2) * Wc + (W - k)]; fb.y = -fb.y;
    }
    const size_t i = ((((size_t)pp * G + g) * (RC / 2) + mg) * NT + blk * CG + c) * 2 + m2;
    dst[i] = make_float2(fa.x - fb.y, fa.y + fb.x);
  }
}
#endif
static __global__ void k_packz_dq(const float* __restrict__ src, float* __restrict__ dst, int C, PackGeom q) {
  const int H = q.H, W = q.W, RC = q.hRC;
  const int NT = CG * (H / RC), G = W / CG, Wc = W / 2 + 1;
  const size_t total = (size_t)C * G * H * CG;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int lane = (int)(i % 4);                     // dst[tile][m / 4][task][m % 4]
  size_t r = i / 4;
  const int task = (int)(r % NT); r /= NT;
  const int mg = (int)(r % (RC / 4)); r /= (RC / 4);
  const int g = (int)(r % G);
  const int ch = (int)(r / G);
  const int m = mg * 4 + lane, c = task % CG, blk = task / CG;
  const int sc = g * CG + c;
  const int h = freq_of_pos_rt(blk * RC + m, H, q.hRA, q.hRB);
  const int k = freq_of_pos_rt((sc % (W / q.wRC)) * q.wRC + sc / (W / q.wRC), W, q.wRA, q.wRB);
  dst[i] = k < Wc ? src[((size_t)ch * H + h) * Wc + k] : src[((size_t)ch * H + (H - h) % H) * Wc + (W - k)];
}

}  // namespace fused
}  // namespace dpx
