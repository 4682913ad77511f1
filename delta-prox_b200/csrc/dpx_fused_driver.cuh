// dpx_fused_driver.cuh — size dispatch + launch sequencing of the fused engine, shared between the CUDA
// engine (dpx_fused_fft.cu) and the CPU emulator library (tests/emu/emu_fused.cpp).  A `Backend` provides
//   row<TW,MODE>(grid, smem_bytes, RowParams), col<TH>(grid, smem_bytes, ColParams),
//   pack<TH,V>(src, dst, planes, H, W, G, zero).
#pragma once
#include <math.h>

#include <type_traits>
#include <vector>

#include "dpx_fused_kernels.cuh"

namespace dpx {
namespace fused {

template <int N, int COLS> struct TileFor;
#define DPX_TILE_FOR(N, A, B, C) \
  template <int COLS> struct TileFor<N, COLS> { using type = fft::Tile<N, A, B, C, COLS>; };
DPX_TILE_FOR(64, 4, 4, 4)
DPX_TILE_FOR(128, 8, 4, 4)
DPX_TILE_FOR(256, 8, 8, 4)
DPX_TILE_FOR(512, 8, 8, 8)
DPX_TILE_FOR(1024, 16, 8, 8)
DPX_TILE_FOR(2048, 16, 16, 8)
DPX_TILE_FOR(4096, 16, 16, 16)
DPX_TILE_FOR(192, 12, 4, 4)        // 3 * 2^k sides: radix-12 first pass
DPX_TILE_FOR(384, 12, 8, 4)
DPX_TILE_FOR(768, 12, 8, 8)
DPX_TILE_FOR(1536, 12, 16, 8)
DPX_TILE_FOR(3072, 12, 16, 16)
DPX_TILE_FOR(320, 10, 8, 4)        // 5 * 2^k sides: radix-10 / radix-20 first pass
DPX_TILE_FOR(640, 10, 8, 8)
DPX_TILE_FOR(1280, 20, 8, 8)
DPX_TILE_FOR(2560, 20, 16, 8)
DPX_TILE_FOR(960, 12, 10, 8)       // widths and heights of the camera formats: 1920 x 1080, 1280 x 720, 2560 x 1440, 1600 x 1200,
DPX_TILE_FOR(1600, 20, 10, 8)      // 1920 x 1200, 3840 x 2160.  Radices 10 / 12 in the second pass; 9 and 15 (odd) only in the
DPX_TILE_FOR(1920, 20, 12, 8)      // column transform: a row transform additionally needs (W / RC) % 4 == 0 (k_rowz)
DPX_TILE_FOR(3840, 20, 12, 16)
DPX_TILE_FOR(720, 10, 9, 8)        // column lengths only
DPX_TILE_FOR(1080, 15, 9, 8)
DPX_TILE_FOR(1200, 15, 10, 8)
DPX_TILE_FOR(1440, 15, 12, 8)
DPX_TILE_FOR(2160, 15, 9, 16)
DPX_TILE_FOR(160, 10, 4, 4)        // further products of the same radices (second pass 10 / 12 next to a last pass of 4 or 16):
DPX_TILE_FOR(400, 10, 10, 4)       // VGA / SVGA sides 480, 600 (column only), 800, 1152 and the 2^k * {9, 25} sides around them
DPX_TILE_FOR(480, 12, 10, 4)
DPX_TILE_FOR(576, 12, 12, 4)
DPX_TILE_FOR(800, 10, 10, 8)       // radix 10 throughout: three co-resident row CTAs (85 registers) without the radix-20 spills
DPX_TILE_FOR(1152, 12, 12, 8)
DPX_TILE_FOR(2304, 12, 12, 16)
DPX_TILE_FOR(3200, 20, 10, 16)
DPX_TILE_FOR(600, 15, 10, 4)       // column lengths only (odd radix 15 / 9 passes)
DPX_TILE_FOR(864, 12, 9, 8)
DPX_TILE_FOR(2400, 15, 10, 16)
DPX_TILE_FOR(2880, 20, 9, 16)
#undef DPX_TILE_FOR

// sizes usable as a row length (W) and as a column length (H) / as a column length only
#ifdef DPX_EXP_SIZES                                   // experiment builds: only the headline size (compile time)
#define DPX_W_SIZES(X) X(2048)
#define DPX_H_ONLY_SIZES(X)
#elif defined(DPX_EMU)                                 // CPU emulator build (tests/emu): the sizes its tests run, one per radix family
#define DPX_W_SIZES(X) X(64) X(128) X(256) X(1024) X(192) X(384) X(320) X(640) X(1280) X(960) X(1920) X(2560) X(480)
#define DPX_H_ONLY_SIZES(X) X(1080) X(4096)            // 2560-point rows / 4096-point columns: the 512-thread, one-CTA-per-SM tiles
#else
#define DPX_W_SIZES(X)                                                                                                          \
  X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(192) X(384) X(768) X(1536) X(3072) X(320) X(640) X(1280) X(2560) X(960) \
  X(1600) X(1920) X(3840) X(160) X(400) X(480) X(576) X(800) X(1152) X(2304) X(3200)
#define DPX_H_ONLY_SIZES(X) X(720) X(1080) X(1200) X(1440) X(2160) X(600) X(864) X(2400) X(2880)
#endif
#define DPX_CASE_TRUE(N) case N:
#define DPX_CASE_CALL(N) case N: f(std::integral_constant<int, N>{}); return true;

inline bool size_supported(int n) {                    // as a row length (and column length)
  switch (n) { DPX_W_SIZES(DPX_CASE_TRUE) return true; default: return false; }
}
inline bool size_supported_h(int n) {                  // as a column length
  switch (n) { DPX_W_SIZES(DPX_CASE_TRUE) DPX_H_ONLY_SIZES(DPX_CASE_TRUE) return true; default: return false; }
}

// calls f(std::integral_constant<int, N>{}) for the matching supported N; returns false if unsupported
template <class F>
inline bool dispatch_size(int n, F&& f) {
  switch (n) { DPX_W_SIZES(DPX_CASE_CALL) default: return false; }
}
template <class F>
inline bool dispatch_size_h(int n, F&& f) {
  switch (n) { DPX_W_SIZES(DPX_CASE_CALL) DPX_H_ONLY_SIZES(DPX_CASE_CALL) default: return false; }
}

inline size_t s_elems(int P, int H, int W) { return (size_t)P * ((W / 2) / CG + 1) * H * CG; }     // float2 count of S
inline size_t packed_elems(int planes, int H, int W) { return (size_t)planes * ((W / 2) / CG + 1) * H * CG; }

// host-side twiddle records (double precision, rounded once): see fft::TwiddleLayout
template <class T>
inline std::vector<float2> make_twiddle_records() {
  using L = fft::TwiddleLayout<T>;
  std::vector<float2> t(L::TOTAL);
  const double two_pi = 6.283185307179586476925286766559;
  for (int j = 0; j < T::MA; ++j)
    for (int q = 0; q < T::RA; ++q) {
      const double a = two_pi * (double)j * q / T::N;
      t[L::A_OFF + ((q / 2) * T::MA + j) * 2 + (q % 2)] = make_float2((float)cos(a), (float)-sin(a));
    }
  for (int j = 0; j < T::MB; ++j)
    for (int q = 0; q < T::RB; ++q) {
      const double a = two_pi * (double)j * q / T::MA;
      t[L::B_OFF + ((q / 2) * T::MB + j) * 2 + (q % 2)] = make_float2((float)cos(a), (float)-sin(a));
    }
  return t;
}
inline std::vector<float2> twiddle_records_for(int n) {
  std::vector<float2> out;
  dispatch_size_h(n, [&](auto nn) { out = make_twiddle_records<typename TileFor<decltype(nn)::value, 1>::type>(); });
  return out;
}

template <class Backend>
struct Driver {
  Backend& be;
  explicit Driver(Backend& b) : be(b) {}

  // column kernel: the persistent TMA-pipelined variant where three padded tiles fit shared memory, else k_col
  template <class TH>
  void launch_col(const ColParams& cp, int nb, int groups, int C) {
    if constexpr (TH::N == 1024 || TH::N == 2048) {
      if (be.sm_count() > 0 && cp.dpsp == nullptr) {
        const int n_tiles = nb * groups * C, per_sm = TH::N >= 2048 ? 1 : 2;
        const int ctas = n_tiles < be.sm_count() * per_sm ? n_tiles : be.sm_count() * per_sm;
        be.template col_tma<TH>(dim3(ctas), ColTmaCfg<TH>::BYTES, cp, n_tiles, nb);
        return;
      }
    }
    be.template col<TH>(dim3(nb, groups, C), TH::SMEM_FLOAT2 * sizeof(float2), cp);
  }

  // packs F(K^T b) (complex, [P,H,Wc]) and sum|OTF|^2 (real, [Cd,H,Wc]) into k_col's record layout
  void pack_constants(int P, int Cd, int H, int W, const float2* fb_std, float2* fbp, const float* dq_std, float* dqp,
                      int C = 0, const float* dpsi_std = nullptr, float* dpsp = nullptr) {
    const int G = (W / 2) / CG;
    dispatch_size_h(H, [&](auto hn) {
      using TH = typename TileFor<decltype(hn)::value, CG>::type;
      if (fb_std) be.template pack<TH, float2>(fb_std, fbp, P, H, W, G, make_float2(0.f, 0.f));
      if (dq_std) be.template pack<TH, float>(dq_std, dqp, Cd, H, W, G, 0.f);
      if (dpsi_std) be.template pack<TH, float>(dpsi_std, dpsp, C, H, W, G, 0.f);
    });
  }

  // n_iters iterations of ADMM (hqs=0) / HQS (hqs=1); state in psi.t[i].v/.u and x; schedules indexed from it0
  void iterate(int B, int C, int H, int W, float2* S, const PsiPack& psi, int hqs, float* x, const float2* fbp,
               const float* dqp, int dq_batch, float wid, float eps, const float* rho, int rho_stride, int it0,
               int n_iters, const float2* tw_h, const float2* tw_w) {
    if (n_iters <= 0) return;
    const int P = B * C, G = (W / 2) / CG;
    dispatch_size(W, [&](auto wn) {
      dispatch_size_h(H, [&](auto hn) {
        using TW = typename TileFor<decltype(wn)::value, ROWS / 2>::type;
        using TH = typename TileFor<decltype(hn)::value, CG>::type;
        RowParams rp;
        rp.C = C; rp.H = H; rp.S = S; rp.psi = psi; rp.hqs = hqs; rp.it = it0; rp.x = x; rp.tw = tw_w;
        ColParams cp;
        cp.C = C; cp.W = W; cp.groups = G + 1; cp.bmul = 1; cp.eps_im = 0.f; cp.dpsp = nullptr; cp.S = S; cp.fbp = fbp; cp.dqp = dqp; cp.dq_batch = dq_batch;
        cp.wid = wid; cp.eps = eps; cp.inv_n = 1.0f / (float)((double)H * W);
        cp.rho.p = rho; cp.rho.stride = rho_stride; cp.rho.it = it0; cp.tw = tw_h;
        const dim3 rgrid(H / ROWS, P);
        const size_t rsm = RowSmem<TW>::BYTES;
        auto row = [&](auto mode) {
          constexpr int MODE = decltype(mode)::value;
          // the once-per-solve first pass only exists in its general (accumulating) form: fewer kernels to compile
          if constexpr (MODE == ROW_FIRST) be.template row<TW, MODE, false>(rgrid, rsm, rp);
          else if (psi.n == 1) be.template row<TW, MODE, true>(rgrid, rsm, rp);
          else be.template row<TW, MODE, false>(rgrid, rsm, rp);
        };
        row(std::integral_constant<int, ROW_FIRST>{});
        for (int k = 0; k < n_iters; ++k) {
          cp.rho.it = it0 + k;
          launch_col<TH>(cp, B, G + 1, C);
          rp.it = it0 + k;
          if (k + 1 < n_iters) {
            if (psi.n == 1 && be.persistent_ctas() > 0)
              be.template row_persist<TW>(dim3(be.persistent_ctas()), RowPersistSmem<TW>::BYTES, rp, (H / ROWS) * P);
            else row(std::integral_constant<int, ROW_MID>{});
          }
          else row(std::integral_constant<int, ROW_LAST>{});
        }
      });
    });
  }

  // ---- plane-pair engine (k_rowz / k_col on Z = X_A + i X_B; see dpx_fused_kernels.cuh) ---------------------------------
  // usable when the two planes of every pair share all solve coefficients: even batch, batch-shared diagonal, shared
  // (not per-sample) rho and lam schedules
  static bool pairs_ok(int B, int dq_batch, int rho_stride, const PsiPack& psi) {
    if (B < 2 || B % 2 != 0 || dq_batch != 1 || rho_stride != 0) return false;
    for (int i = 0; i < psi.n; ++i)
      if (psi.t[i].lam_stride != 0) return false;
    return true;
  }
  static size_t pair_elems(int planes, int H, int W) { return (size_t)planes * H * W; }   // float2 (fbz, S) / float (dqz) count

  void pack_constants_pairs(int B, int C, int H, int W, const float2* fb_std, float2* fbz, const float* dq_std, float* dqz,
                            const float* dpsi_std = nullptr, float* dpsz = nullptr) {
    dispatch_size(W, [&](auto wn) {
      dispatch_size_h(H, [&](auto hn) {
        using TW = typename TileFor<decltype(wn)::value, ROWS>::type;
        using TH = typename TileFor<decltype(hn)::value, CG>::type;
        const PackGeom q{H, W, TH::RA, TH::RB, TH::RC, TW::RA, TW::RB, TW::RC};
        if (fb_std) be.packz_fb(fb_std, fbz, (B / 2) * C, C, q);
        if (dq_std) be.packz_dq(dq_std, dqz, C, q);
        if (dpsi_std) be.packz_dq(dpsi_std, dpsz, C, q);
      });
    });
  }

  void iterate_pairs(int B, int C, int H, int W, float2* S, const PsiPack& psi, int hqs, float* x, const float2* fbz,
                     const float* dqz, float wid, float eps, const float* rho, int it0, int n_iters, const float2* tw_h,
                     const float2* tw_w) {
    if (n_iters <= 0) return;
    const int PP = (B / 2) * C, G = W / CG;
    dispatch_size(W, [&](auto wn) {
      dispatch_size_h(H, [&](auto hn) {
        using TW = typename TileFor<decltype(wn)::value, ROWS>::type;
        using TH = typename TileFor<decltype(hn)::value, CG>::type;
        RowParams rp;
        rp.C = C; rp.H = H; rp.S = S; rp.psi = psi; rp.hqs = hqs; rp.it = it0; rp.x = x; rp.tw = tw_w;
        ColParams cp;
        cp.C = C; cp.W = W; cp.groups = G; cp.bmul = 2; cp.eps_im = eps; cp.dpsp = nullptr; cp.S = S; cp.fbp = fbz; cp.dqp = dqz; cp.dq_batch = 1;
        cp.wid = wid; cp.eps = eps; cp.inv_n = 1.0f / (float)((double)H * W);
        cp.rho.p = rho; cp.rho.stride = 0; cp.rho.it = it0; cp.tw = tw_h;
        const dim3 rgrid(H / ROWS, PP);
        const size_t rsm = TW::SMEM_FLOAT2 * sizeof(float2);
        auto row = [&](auto mode) {
          constexpr int MODE = decltype(mode)::value;
          if constexpr (MODE == ROW_FIRST) be.template rowz<TW, MODE, false>(rgrid, rsm, rp);
          else if (psi.n == 1) be.template rowz<TW, MODE, true>(rgrid, rsm, rp);
          else be.template rowz<TW, MODE, false>(rgrid, rsm, rp);
        };
        if (psi.n == 1 && be.persistent_ctas() > 0) {
          using TW2 = typename TileFor<decltype(wn)::value, ZR>::type;
          be.template rowz_persist<TW2, PM_FIRST>(dim3(be.persistent_ctas() / 2 * RowZPersistSmem<TW2>::CTAS_PER_SM), RowZPersistSmem<TW2>::BYTES, rp,
                                                  (H / ZR) * PP);
        } else row(std::integral_constant<int, ROW_FIRST>{});
        for (int k = 0; k < n_iters; ++k) {
          cp.rho.it = it0 + k;
          launch_col<TH>(cp, B / 2, G, C);
          rp.it = it0 + k;
          if (psi.n == 1 && be.persistent_ctas() > 0) {
            using TW2 = typename TileFor<decltype(wn)::value, ZR>::type;
            const dim3 pg(be.persistent_ctas() / 2 * RowZPersistSmem<TW2>::CTAS_PER_SM);
            if (k + 1 < n_iters) be.template rowz_persist<TW2, PM_MID>(pg, RowZPersistSmem<TW2>::BYTES, rp, (H / ZR) * PP);
            else be.template rowz_persist<TW2, PM_LAST>(pg, RowZPersistSmem<TW2>::BYTES, rp, (H / ZR) * PP);   // last: x, v, u out
          }
          else if (k + 1 < n_iters) row(std::integral_constant<int, ROW_MID>{});
          else row(std::integral_constant<int, ROW_LAST>{});
        }
      });
    });
  }

  // ---- staged x-update (an external prox sits between the stages): x <- closed-form solve of the current state --------
  //   rows: t = sum_i s_i (v_i - u_i) -> forward row FFT;  columns: FFT, solve, inverse FFT;  rows: inverse FFT -> x
  void xupdate(bool pairs, int B, int C, int H, int W, float2* S, const PsiPack& psi, int hqs, float* x, const float2* fbp,
               const float* dqp, int dq_batch, float wid, float eps, const float* rho, int rho_stride, int it,
               const float2* tw_h, const float2* tw_w, const float* dpsp = nullptr) {
    const int P = B * C;
    dispatch_size(W, [&](auto wn) {
      dispatch_size_h(H, [&](auto hn) {
        using TH = typename TileFor<decltype(hn)::value, CG>::type;
        RowParams rp;
        rp.C = C; rp.H = H; rp.S = S; rp.psi = psi; rp.hqs = hqs; rp.it = it; rp.x = x; rp.tw = tw_w;
        ColParams cp;
        cp.C = C; cp.W = W; cp.S = S; cp.fbp = fbp; cp.dqp = dqp; cp.dpsp = dpsp; cp.wid = wid; cp.eps = eps;
        cp.inv_n = 1.0f / (float)((double)H * W);
        cp.rho.p = rho; cp.rho.it = it; cp.tw = tw_h;
        if (pairs) {
          using TW = typename TileFor<decltype(wn)::value, ROWS>::type;
          const int G = W / CG;
          cp.groups = G; cp.bmul = 2; cp.eps_im = eps; cp.dq_batch = 1; cp.rho.stride = 0;
          const dim3 rgrid(H / ROWS, P / 2);
          const size_t rsm = TW::SMEM_FLOAT2 * sizeof(float2);
          using TW2 = typename TileFor<decltype(wn)::value, ZR>::type;
          const bool persist = be.persistent_ctas() > 0;
          const dim3 pg(persist ? be.persistent_ctas() / 2 * RowZPersistSmem<TW2>::CTAS_PER_SM : 1);
          if (persist && psi.n == 1) be.template rowz_persist<TW2, PM_FIRST>(pg, RowZPersistSmem<TW2>::BYTES, rp, (H / ZR) * (P / 2));
          else be.template rowz<TW, ROW_FIRST, false>(rgrid, rsm, rp);
          launch_col<TH>(cp, B / 2, G, C);
          if (persist) be.template rowz_persist<TW2, PM_XONLY>(pg, RowZPersistSmem<TW2>::BYTES, rp, (H / ZR) * (P / 2));
          else be.template rowz<TW, ROW_XONLY, false>(rgrid, rsm, rp);
        } else {
          using TW = typename TileFor<decltype(wn)::value, ROWS / 2>::type;
          const int G = (W / 2) / CG;
          cp.groups = G + 1; cp.bmul = 1; cp.eps_im = 0.f; cp.dq_batch = dq_batch; cp.rho.stride = rho_stride;
          const dim3 rgrid(H / ROWS, P);
          const size_t rsm = RowSmem<TW>::BYTES;
          be.template row<TW, ROW_FIRST, false>(rgrid, rsm, rp);
          launch_col<TH>(cp, B, G + 1, C);
          be.template row<TW, ROW_XONLY, false>(rgrid, rsm, rp);
        }
      });
    });
  }
};

}  // namespace fused
}  // namespace dpx
