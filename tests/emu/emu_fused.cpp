// emu_fused.cpp — TEST INFRASTRUCTURE: runs the fused FFT engine's kernels (same source as the GPU build)
// on the CPU through tests/emu/cuda_emu.h.  Built by tests/test_fused_emulator.py with g++ and driven via ctypes.
#define DPX_EMU
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "dpx_fused_driver.cuh"

using namespace dpx;
using namespace dpx::fused;

namespace {
struct EmuBackend {
  template <class TW, int MODE, bool SINGLE>
  void row(dim3 grid, size_t smem, RowParams p) {
    emu::launch(grid, dim3(kThreads), smem, [=]() { k_row<TW, MODE, SINGLE>(p); });
  }
  int persist = 5;                                    // emulated persistent grid (forces every CTA to loop)
  int persistent_ctas() const { return persist; }
  template <class TW>
  void row_persist(dim3 grid, size_t smem, RowParams p, int n_tiles) {
    emu::launch(grid, dim3(kThreads), smem, [=]() { k_row_mid_persist<TW>(p, n_tiles); });
  }
  int sms = 3;                                        // emulated SM count of the persistent TMA column kernel
  int sm_count() const { return sms; }
  template <class TH>
  void col_tma(dim3 grid, size_t smem, ColParams p, int n_tiles, int nb) {
    emu::launch(grid, dim3(ColTmaCfg<TH>::NT), smem, [=]() { k_col_tma<TH>(p, n_tiles, nb); });
  }
  template <class TH>
  void col(dim3 grid, size_t smem, ColParams p) {
    emu::launch(grid, dim3(ColThreads<TH>::value), smem, [=]() { k_col<TH>(p); });
  }
  template <class TW, int MODE, bool SINGLE>
  void rowz(dim3 grid, size_t smem, RowParams p) {
    emu::launch(grid, dim3(kThreads), smem, [=]() { k_rowz<TW, MODE, SINGLE>(p); });
  }
  template <class TW, int PM>
  void rowz_persist(dim3 grid, size_t smem, RowParams p, int n_tiles) {
    emu::launch(grid, dim3(RowZPersistSmem<TW>::THREADS), smem, [=]() { k_rowz_mid_persist<TW, PM>(p, n_tiles); });
  }
  void packz_fb(const float2* src, float2* dst, int pairs, int C, PackGeom q) {
    const size_t total = (size_t)pairs * q.H * q.W;
    emu::launch(dim3((unsigned)((total + 255) / 256)), dim3(256), 0, [=]() { k_packz_fb(src, dst, pairs, C, q); });
  }
  void packz_dq(const float* src, float* dst, int C, PackGeom q) {
    const size_t total = (size_t)C * q.H * q.W;
    emu::launch(dim3((unsigned)((total + 255) / 256)), dim3(256), 0, [=]() { k_packz_dq(src, dst, C, q); });
  }
  template <class TH, typename V>
  void pack(const V* src, V* dst, int planes, int H, int W, int G, V zero) {
    const size_t total = (size_t)planes * (G + 1) * H * CG;
    emu::launch(dim3((unsigned)((total + 255) / 256)), dim3(256), 0, [=]() { k_pack<TH, V>(src, dst, planes, H, W, G, zero); });
  }
};
}  // namespace

extern "C" __attribute__((visibility("default"))) int emu_fused_supported(int n) { return size_supported(n) ? 1 : 0; }
extern "C" __attribute__((visibility("default"))) int emu_fused_supported_h(int n) { return size_supported_h(n) ? 1 : 0; }

// x, v[i], u[i]: [B,C,H,W] fp32 (updated in place); fb_std: complex64 [B*C,H,W/2+1]; dq_std: [Cd,H,W/2+1], Cd = C or B*C
// psi_*: per-term arrays; lam: [n_psi][T]; rho: [T]; offsets off[i] may be null
extern "C" __attribute__((visibility("default"))) int emu_fused_run(int B, int C, int H, int W, int n_psi, const int* prox, const float* scale, const float* alpha,
                             const float* beta, float* x, float** v, float** u, const float** off, const float* fb_std,
                             const float* dq_std, int dq_batch, float wid, float eps, const float* rho, const float* lam, int T,
                             int hqs) {
  if (!size_supported_h(H) || !size_supported(W)) return 1;
  const int P = B * C, Cd = dq_batch > 1 ? P : C;
  std::vector<float2> S(s_elems(P, H, W), make_float2(0.f, 0.f));
  std::vector<float2> fbp(packed_elems(P, H, W));
  std::vector<float> dqp(std::max(packed_elems(Cd, H, W), (size_t)C * H * W));     // the pair layout holds the full spectrum
  auto twh = twiddle_records_for(H), tww = twiddle_records_for(W);
  EmuBackend be;
  Driver<EmuBackend> drv(be);
  PsiPack pk;
  pk.n = n_psi;
  for (int i = 0; i < n_psi; ++i) {
    PsiTerm& t = pk.t[i];
    t.prox = prox[i]; t.linop = 0; t.scale = scale[i]; t.alpha = alpha[i]; t.beta = beta[i]; t.inv_beta = 1.f / beta[i];
    t.lo = 0.f; t.hi = 1.f; t.v = v[i]; t.u = hqs ? nullptr : u[i]; t.off = off ? off[i] : nullptr;
    t.lam = lam + (size_t)i * T; t.lam_stride = 0;
  }
  const bool pairs = !getenv("DPX_EMU_NO_PAIRS") && Driver<EmuBackend>::pairs_ok(B, dq_batch, 0, pk);
  // DPX_EMU_LINOPS="1,2": the solve's denominator gains the diagonal of stencil-gradient psi linops (1 = along H, 2 = along
  // W), sum_i |F(K_i)|^2 = 2 - 2 cos(2 pi k / n), built here and packed like the plan does (their right-hand side
  // sum_i K_i^T (v_i - u_i) is formed by the stencil kernel and handed to the fused engine as one identity term)
  std::vector<float> dpsi_std, dpsp;
  if (const char* lo = getenv("DPX_EMU_LINOPS")) {
    const int Wc = W / 2 + 1;
    dpsi_std.assign((size_t)C * H * Wc, 0.f);
    int i = 0;
    for (const char* q = lo; *q; ++q) {
      if (*q == ',') continue;
      const int l = *q - '0';
      const double two_pi = 6.283185307179586476925286766559;
      if (l == 1 || l == 2)
        for (int c = 0; c < C; ++c)
          for (int h = 0; h < H; ++h)
            for (int k = 0; k < Wc; ++k)
              dpsi_std[((size_t)c * H + h) * Wc + k] += (float)(2.0 - 2.0 * cos(two_pi * (l == 1 ? (double)h / H : (double)k / W)));
      ++i;
    }
    dpsp.assign(std::max(packed_elems(C, H, W), (size_t)C * H * W), 0.f);
  }
  if (getenv("DPX_EMU_XUPDATE")) {    // staged form: only the x-update of iteration 0 (rows: rhs + FFT, columns, rows: inverse -> x)
    const float* ds = dpsi_std.empty() ? nullptr : dpsi_std.data();
    if (pairs) drv.pack_constants_pairs(B, C, H, W, reinterpret_cast<const float2*>(fb_std), fbp.data(), dq_std, dqp.data(), ds, dpsp.data());
    else drv.pack_constants(P, Cd, H, W, reinterpret_cast<const float2*>(fb_std), fbp.data(), dq_std, dqp.data(), C, ds, dpsp.data());
    drv.xupdate(pairs, B, C, H, W, S.data(), pk, hqs, x, fbp.data(), dqp.data(), dq_batch, wid, eps, rho, 0, 0, twh.data(), tww.data(),
                ds ? dpsp.data() : nullptr);
    return pairs ? 2 : 0;
  }
  if (pairs) {      // plane-pair engine: what the CUDA engine selects for an even batch with shared schedules
    drv.pack_constants_pairs(B, C, H, W, reinterpret_cast<const float2*>(fb_std), fbp.data(), dq_std, dqp.data());
    drv.iterate_pairs(B, C, H, W, S.data(), pk, hqs, x, fbp.data(), dqp.data(), wid, eps, rho, 0, T, twh.data(), tww.data());
    return 2;
  }
  drv.pack_constants(P, Cd, H, W, reinterpret_cast<const float2*>(fb_std), fbp.data(), dq_std, dqp.data());
  drv.iterate(B, C, H, W, S.data(), pk, hqs, x, fbp.data(), dqp.data(), dq_batch, wid, eps, rho, 0, 0, T, twh.data(), tww.data());
  return 0;
}
