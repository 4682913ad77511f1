// dpx_fused_fft.cu — the fused sm_100a FFT engine: two kernels per ADMM/HQS iteration
// (dpx_fused_kernels.cuh) instead of cuFFT R2C + solve + cuFFT C2R + prox/dual.
//
// Generic transforms (constant hoisting, stand-alone spectral filters, algorithms the fused kernels do not
// cover) are delegated to an embedded cuFFT engine, so this engine is a strict superset of it.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <new>
#include <vector>

#include "dpx_fft.cuh"
#include <cuda.h>

#include "dpx_fused_launch.cuh"

namespace dpx {

int make_cufft_engine(const Geom& g, FftEngine** out);   // dpx_fft.cu

namespace {

using namespace fused;

struct CudaBackend {
  cudaStream_t s;
  int rc = DPX_OK;

  void done(cudaError_t e) {
    ++g_launches;
    if (e != cudaSuccess && rc == DPX_OK) { set_error("fused kernel launch failed: %s", cudaGetErrorString(e)); rc = DPX_ERR_CUDA; }
  }
  void after() { done(cudaGetLastError()); }
  template <class TW, int MODE, bool SINGLE>
  void row(dim3 grid, size_t smem, const RowParams& p) {
    if (rc) return;
    RowParams q = p;
    q.pdl = pdl;
    done(launch::row<TW, MODE, SINGLE>(grid, smem, q, s));
  }
  int n_persist = 0;                                  // CTAs of the persistent row kernel (2 per SM); 0 disables it
  int persistent_ctas() const { return n_persist; }
  template <class TW>
  void row_persist(dim3 grid, size_t smem, const RowParams& p, int n_tiles) {
    if (rc) return;
    RowParams q = p;
    q.pdl = pdl;
    done(launch::row_persist<TW>(grid, smem, q, n_tiles, s));
  }
  int n_sm = 0;                                       // SMs for the persistent TMA column kernel; 0 disables it
  int sm_count() const { return n_sm; }
  template <class TH>
  void col_tma(dim3 grid, size_t smem, const ColParams& p, int n_tiles, int nb) {
    if (rc) return;
    done(launch::col_tma<TH>(grid, smem, p, n_tiles, nb, s));
  }
  static constexpr int kRowCtrs = 4096;                 // dynamic-tile counters: one per persistent row launch of a call (zeroed per call)
  int* row_ctr = nullptr;
  int row_ctr_next = 0, row_dyn = 1;
  unsigned long long *trace_col = nullptr, *trace_row = nullptr;
  const void* row_smap = nullptr;                     // tensor map of S for the persistent pair kernel's staging (DPX_ROW_TMA=0: LDGSTS)
  int pdl = 0;                                        // k_col / k_rowz_mid_persist as programmatic dependent launches (DPX_PDL=0: off)
  int col_bulk = 0;                                   // column tile staged by TMA bulk copies (DPX_COL_BULK=0: LDGSTS)
  template <class TH>
  void col(dim3 grid, size_t smem, const ColParams& p) {
    if (rc) return;
    ColParams q = p;
    q.trace = trace_col;
    q.bulk = col_bulk;
    q.pdl = pdl;
    done(launch::col<TH>(grid, smem + (col_bulk ? 16 : 0), q, s));
  }
  template <class TW, int MODE, bool SINGLE>
  void rowz(dim3 grid, size_t smem, const RowParams& p) {
    if (rc) return;
    RowParams q = p;
    q.pdl = pdl;
    done(launch::rowz<TW, MODE, SINGLE>(grid, smem, q, s));
  }
  template <class TW, int PM = PM_MID>
  void rowz_persist(dim3 grid, size_t smem, const RowParams& p, int n_tiles) {
    if (rc) return;
    RowParams q = p;
    q.trace = trace_row;
    q.smap = row_smap;
    q.pdl = pdl;
    if (row_ctr && row_dyn) { q.ctr = row_ctr + (row_ctr_next++ % kRowCtrs); }
    done(launch::rowz_persist<TW, PM>(grid, smem, q, n_tiles, s));
  }
  void packz_fb(const float2* src, float2* dst, int pairs, int C, const PackGeom& q) {
    if (rc) return;
    const size_t smem = (size_t)8 * (q.W / 2 + 1) * sizeof(float2);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_packz_fb_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
    k_packz_fb_rows<<<dim3(q.hRC / 2, q.H / q.hRC, pairs), 256, smem, s>>>(src, dst, pairs, C, q);
    after();
  }
  void packz_dq(const float* src, float* dst, int C, const PackGeom& q) {
    if (rc) return;
    const size_t total = (size_t)C * q.H * q.W;
    k_packz_dq<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(src, dst, C, q);
    after();
  }
  template <class TH, typename V>
  void pack(const V* src, V* dst, int planes, int H, int W, int G, V zero) {
    if (rc) return;
    done(launch::pack<TH, V>(src, dst, planes, H, W, G, zero, s));
  }
};

class FusedEngine final : public FftEngine {
 public:
  int init(const Geom& g) {
    g_ = g;
    int rc = make_cufft_engine(g, &inner_);
    if (rc) return rc;
    const size_t ns = s_elems(g.P, g.H, g.W);
    DPX_CUDA(cudaMalloc(&S_, ns * sizeof(float2)));
    DPX_CUDA(cudaMemset(S_, 0, ns * sizeof(float2)));
    DPX_CUDA(cudaMalloc(&fbp_, ns * sizeof(float2)));
    DPX_CUDA(cudaMemset(fbp_, 0, ns * sizeof(float2)));
    bytes_ = 2 * ns * sizeof(float2);
    {
      int dev = 0, sms = 0;
      DPX_CUDA(cudaGetDevice(&dev));
      DPX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      const char* env = getenv("DPX_ROW_PERSIST");            // 0 disables the persistent row kernel (for A/B runs)
      persist_ctas_ = (env && env[0] == '0') ? 0 : 2 * sms;
      const char* te = getenv("DPX_COL_TMA");                 // 1 selects the persistent TMA column kernel
      col_tma_sms_ = (te && te[0] == '1') ? sms : 0;          // opt-in: measured slower than 3 co-resident k_col CTAs (profiles/README.md)
      const char* pe = getenv("DPX_PAIRS");                   // 0 disables the plane-pair engine (for A/B runs)
      pairs_enabled_ = !(pe && pe[0] == '0');
      DPX_CUDA(cudaMalloc(&row_ctr_, sizeof(int) * CudaBackend::kRowCtrs));
      const char* be_ = getenv("DPX_COL_BULK");
      col_bulk_ = (be_ && be_[0] == '0') ? 0 : 1;             // 0 = LDGSTS staging of the column tile (A/B runs)
      const char* te2 = getenv("DPX_ROW_TMA");
      row_tma_ = (te2 && te2[0] == '0') ? 0 : 1;             // 0 = LDGSTS staging of the spectrum rows (A/B runs)
      const char* de = getenv("DPX_ROW_DYN");                 // 0 = static round-robin tiles in the persistent pair kernel
      row_dyn_ = !(de && de[0] == '0');
      // k_col / k_rowz_mid_persist as programmatic dependent launches (griddep_wait in the kernels): the next kernel's launch, CTA
      // scheduling and barrier set-up overlap the previous kernel's tail.  Neutral at 2048^2 (580 us per iteration either way), -23 % per
      // iteration where a kernel is a single short wave (2 x [3,256,256]: 13.3 -> 10.2 us).  DPX_PDL=0 restores plain launches (A/B runs)
      const char* pd = getenv("DPX_PDL");
      pdl_ = (pd && pd[0] == '0') ? 0 : 1;
      trace_path_ = getenv("DPX_TRACE");
      if (trace_path_) {
        DPX_CUDA(cudaMalloc(&trace_col_, kTraceRecs * 16 * sizeof(unsigned long long)));
        DPX_CUDA(cudaMalloc(&trace_row_, kTraceRecs * 16 * sizeof(unsigned long long)));
        DPX_CUDA(cudaMemset(trace_col_, 0, kTraceRecs * 16 * sizeof(unsigned long long)));
        DPX_CUDA(cudaMemset(trace_row_, 0, kTraceRecs * 16 * sizeof(unsigned long long)));
      }
    }
    rc = upload_twiddles(g.H, &tw_h_);
    if (!rc) rc = upload_twiddles(g.W, &tw_w_);
    return rc;
  }
  int r2c(const float* in, float2* out, cudaStream_t s) override { return inner_->r2c(in, out, s); }
  int c2r(float2* in, float* out, cudaStream_t s) override { return inner_->c2r(in, out, s); }
  size_t workspace_bytes() const override { return bytes_ + (inner_ ? inner_->workspace_bytes() : 0); }
  void dump_trace() {
    if (!trace_path_ || !trace_col_) return;
    std::vector<unsigned long long> h(2 * kTraceRecs * 16);
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), trace_col_, kTraceRecs * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(h.data() + kTraceRecs * 16, trace_row_, kTraceRecs * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(trace_path_, "wb")) { fwrite(h.data(), sizeof(unsigned long long), h.size(), f); fclose(f); }
    cudaFree(trace_col_); cudaFree(trace_row_); trace_col_ = trace_row_ = nullptr;
  }
  void destroy() override {
    dump_trace();
    if (inner_) inner_->destroy();
    cudaFree(S_); cudaFree(fbp_); cudaFree(dqp_); cudaFree(dpsp_); cudaFree(S1_); cudaFree(fbp1_); cudaFree(dqp1_);
    cudaFree(tw_h_); cudaFree(tw_w_); cudaFree(row_ctr_); cudaFree(smap_dev_);
    if (side_) { cudaStreamDestroy(side_); cudaEventDestroy(ev_fork_); cudaEventDestroy(ev_join_); }
    delete this;
  }
  bool fused() const override { return true; }
  int engine_mode() const override {
    return packed_mode_ == PAIRS ? DPX_ENGINE_FUSED_PAIRS : (packed_mode_ == FLAT ? DPX_ENGINE_FUSED_FLAT : DPX_ENGINE_FUSED_PLANES);
  }
  void reset_constants() override { dq_set_ = false; dq_std_ = nullptr; dq_dirty_ = true; }
  void set_dpsi(const float* dpsi_std) override { dpsi_std_ = dpsi_std; dpsi_dirty_ = true; }

  // Constants are packed lazily, in the layout of the engine variant the next fused_iters call selects (half-spectrum
  // planes or plane pairs): set_constants only records which standard-layout arrays (owned by the plan) changed.
  int set_constants(const float2* fb_std, const float* dq_std, int dq_batch, cudaStream_t s) override {
    fb_std_ = fb_std; fb_dirty_ = true;                 // nullptr = zero right-hand side
    if (dq_std || !dq_set_) { dq_std_ = dq_std; dq_dirty_ = true; dq_set_ = dq_set_ || dq_std != nullptr; }
    dq_batch_ = dq_batch;
    (void)s;
    return DPX_OK;
  }

  // FLAT mode runs the odd last plane on a side stream, concurrently with the pairs (small grids: they fill each other's tails)
  int fork_side(cudaStream_t s) {
    if (!side_) {
      DPX_CUDA(cudaStreamCreateWithFlags(&side_, cudaStreamNonBlocking));
      DPX_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
      DPX_CUDA(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
    }
    DPX_CUDA(cudaEventRecord(ev_fork_, s));
    DPX_CUDA(cudaStreamWaitEvent(side_, ev_fork_, 0));
    return DPX_OK;
  }
  int join_side(cudaStream_t s) {
    DPX_CUDA(cudaEventRecord(ev_join_, side_));
    DPX_CUDA(cudaStreamWaitEvent(s, ev_join_, 0));
    return DPX_OK;
  }

  void set_channel_shared(bool shared) override { if (shared != channel_shared_) { channel_shared_ = shared; fb_dirty_ = dq_dirty_ = true; } }

  enum Mode { PLANES = 0, PAIRS = 1, FLAT = 2 };
  // PLANES: half-spectrum engine.  PAIRS: planes (b, ch) + i (b+1, ch), even batch.  FLAT: the diagonal is shared by all
  // channels (grey PSF), so ANY two planes pair: planes (2k, 2k+1) of the flat plane list ride the pair engine as a
  // (B' = P_even, C' = 1) problem and an odd last plane goes through the half-spectrum engine on its own (B' = C' = 1) --
  // this is what a single RGB image (B = 1, C = 3) gets.
  Mode select_mode(const PsiPack& psi, int rho_stride) const {
    if (!pairs_enabled_) return PLANES;
    if (Driver<CudaBackend>::pairs_ok(g_.B, dq_batch_, rho_stride, psi)) return PAIRS;
    if (channel_shared_ && !dpsi_std_ && g_.P >= 2 && Driver<CudaBackend>::pairs_ok(2, dq_batch_, rho_stride, psi)) return FLAT;
    return PLANES;
  }

  // packs whatever changed, in the layout of the engine variant selected for this call
  int prepare(const PsiPack& psi, int rho_stride, cudaStream_t s, CudaBackend& be, Mode* mode_out) {
    Driver<CudaBackend> drv(be);
    const Mode mode = select_mode(psi, rho_stride);
    if (mode != packed_mode_) { fb_dirty_ = dq_dirty_ = dpsi_dirty_ = true; packed_mode_ = mode; }
    if (dpsi_std_ && !dpsp_) {
      const size_t np = std::max(Driver<CudaBackend>::pair_elems(g_.C, g_.H, g_.W), packed_elems(g_.C, g_.H, g_.W));
      DPX_CUDA(cudaMalloc(&dpsp_, np * sizeof(float)));
      bytes_ += np * sizeof(float);
      dpsi_dirty_ = true;
    }
    const int Cd = dq_batch_ > 1 ? g_.P : g_.C;
    const size_t nd = mode == PLANES ? packed_elems(Cd, g_.H, g_.W) : Driver<CudaBackend>::pair_elems(mode == FLAT ? 1 : g_.C, g_.H, g_.W);
    if (dqp_cap_ < nd) {
      cudaFree(dqp_); dqp_ = nullptr;
      DPX_CUDA(cudaMalloc(&dqp_, nd * sizeof(float)));
      bytes_ += (nd - dqp_cap_) * sizeof(float);
      dqp_cap_ = nd;
      dq_dirty_ = true;
    }
    if (mode == FLAT && (g_.P % 2) && !S1_) {            // the odd last plane: its own spectrum and packed constants
      const size_t n1 = s_elems(1, g_.H, g_.W);
      DPX_CUDA(cudaMalloc(&S1_, n1 * sizeof(float2)));
      DPX_CUDA(cudaMemsetAsync(S1_, 0, n1 * sizeof(float2), s));
      DPX_CUDA(cudaMalloc(&fbp1_, n1 * sizeof(float2)));
      DPX_CUDA(cudaMalloc(&dqp1_, packed_elems(1, g_.H, g_.W) * sizeof(float)));
      bytes_ += 2 * n1 * sizeof(float2) + packed_elems(1, g_.H, g_.W) * sizeof(float);
      fb_dirty_ = dq_dirty_ = true;
    }
    if (dq_dirty_ && !dq_std_) DPX_CUDA(cudaMemsetAsync(dqp_, 0, nd * sizeof(float), s));
    if (fb_dirty_ && !fb_std_) DPX_CUDA(cudaMemsetAsync(fbp_, 0, s_elems(g_.P, g_.H, g_.W) * sizeof(float2), s));
    const float* dps = (dpsi_dirty_ && dpsi_std_) ? dpsi_std_ : nullptr;
    const float2* fbs = fb_dirty_ ? fb_std_ : nullptr;
    const float* dqs = dq_dirty_ ? dq_std_ : nullptr;
    if (mode == PAIRS) drv.pack_constants_pairs(g_.B, g_.C, g_.H, g_.W, fbs, fbp_, dqs, dqp_, dps, dpsp_);
    else if (mode == PLANES) drv.pack_constants(g_.P, Cd, g_.H, g_.W, fbs, fbp_, dqs, dqp_, g_.C, dps, dpsp_);
    else {
      const int Pe = g_.P - g_.P % 2;
      drv.pack_constants_pairs(Pe, 1, g_.H, g_.W, fbs, fbp_, dqs, dqp_);
      if (g_.P % 2) {
        if (fb_dirty_ && !fb_std_) DPX_CUDA(cudaMemsetAsync(fbp1_, 0, s_elems(1, g_.H, g_.W) * sizeof(float2), s));
        if (dq_dirty_ && !dq_std_) DPX_CUDA(cudaMemsetAsync(dqp1_, 0, packed_elems(1, g_.H, g_.W) * sizeof(float), s));
        drv.pack_constants(1, 1, g_.H, g_.W, fbs ? fbs + (size_t)Pe * g_.splane : nullptr, fbp1_, dqs, dqp1_);
      }
    }
    fb_dirty_ = dq_dirty_ = dpsi_dirty_ = false;
    *mode_out = mode;
    return DPX_OK;
  }

  // the last plane of an odd plane count as a (B = C = 1) problem of its own: every per-element pointer shifted to it
  PsiPack last_plane(const PsiPack& psi, float** x) const {
    const size_t off = (size_t)(g_.P - 1) * g_.plane;
    PsiPack q = psi;
    for (int i = 0; i < q.n; ++i) {
      if (q.t[i].v) q.t[i].v += off;
      if (q.t[i].u) q.t[i].u += off;
      if (q.t[i].off) q.t[i].off += off;
    }
    if (*x) *x += off;
    return q;
  }

  int fused_iters(const Geom& g, const PsiPack& psi, bool hqs, float* x, const float2*, const float*, int, float wid,
                  float eps, const float* rho, int rho_stride, int it0, int n_iters, cudaStream_t s) override {
    CudaBackend be{s};
    be.n_persist = persist_ctas_;
    be.n_sm = col_tma_sms_;
    be.trace_col = trace_col_; be.trace_row = trace_row_;
    be.col_bulk = col_bulk_; be.row_ctr = row_ctr_; be.pdl = pdl_;
    be.row_smap = nullptr; be.row_dyn = row_dyn_ && n_iters + 1 <= CudaBackend::kRowCtrs;   // one zeroed counter per persistent launch:
    if (row_ctr_ && be.row_dyn) DPX_CUDA(cudaMemsetAsync(row_ctr_, 0, sizeof(int) * (n_iters + 1), s));   // first pass + n_iters iterations
    Mode mode = PLANES;
    int rc = prepare(psi, rho_stride, s, be, &mode);
    if (rc) return rc;
    Driver<CudaBackend> drv(be);
    if (row_tma_ && mode != PLANES) {
      if (ensure_smap(mode == PAIRS ? (g.B / 2) * g.C : (g.P - g.P % 2) / 2, s) == DPX_OK) be.row_smap = smap_dev_;
      else row_tma_ = 0;                                  // driver without cuTensorMapEncodeTiled: LDGSTS staging
    }
    if (mode == PAIRS)
      drv.iterate_pairs(g.B, g.C, g.H, g.W, S_, psi, hqs ? 1 : 0, x, fbp_, dqp_, wid, eps, rho, it0, n_iters, tw_h_, tw_w_);
    else if (mode == PLANES)
      drv.iterate(g.B, g.C, g.H, g.W, S_, psi, hqs ? 1 : 0, x, fbp_, dqp_, dq_batch_, wid, eps, rho, rho_stride, it0, n_iters,
                  tw_h_, tw_w_);
    else {
      if (g.P % 2) {
        rc = fork_side(s);
        if (rc) return rc;
        CudaBackend be1{side_};
        be1.n_persist = persist_ctas_;
        Driver<CudaBackend> drv1(be1);
        float* x1 = x;
        const PsiPack one = last_plane(psi, &x1);
        drv1.iterate(1, 1, g.H, g.W, S1_, one, hqs ? 1 : 0, x1, fbp1_, dqp1_, 1, wid, eps, rho, 0, it0, n_iters, tw_h_, tw_w_);
        if (be1.rc) return be1.rc;
      }
      drv.iterate_pairs(g.P - g.P % 2, 1, g.H, g.W, S_, psi, hqs ? 1 : 0, x, fbp_, dqp_, wid, eps, rho, it0, n_iters, tw_h_, tw_w_);
      if (g.P % 2) { rc = join_side(s); if (rc) return rc; }
    }
    return be.rc;
  }

  int fused_xupdate(const Geom& g, const PsiPack& psi, bool hqs, float* x, float wid, float eps, const float* rho,
                    int rho_stride, int it, cudaStream_t s) override {
    CudaBackend be{s};
    be.n_sm = col_tma_sms_;
    be.n_persist = persist_ctas_;
    be.col_bulk = col_bulk_; be.row_ctr = row_ctr_; be.pdl = pdl_; be.row_dyn = row_dyn_; be.trace_col = trace_col_; be.trace_row = trace_row_;
    if (row_ctr_ && be.row_dyn) DPX_CUDA(cudaMemsetAsync(row_ctr_, 0, sizeof(int) * 4, s));   // two persistent row launches per x-update
    Mode mode = PLANES;
    int rc = prepare(psi, rho_stride, s, be, &mode);
    if (rc) return rc;
    if (row_tma_ && mode != PLANES) {
      if (ensure_smap(mode == PAIRS ? (g.B / 2) * g.C : (g.P - g.P % 2) / 2, s) == DPX_OK) be.row_smap = smap_dev_;
      else row_tma_ = 0;
    }
    Driver<CudaBackend> drv(be);
    const float* dps = dpsi_std_ ? dpsp_ : nullptr;
    if (mode != FLAT) {
      drv.xupdate(mode == PAIRS, g.B, g.C, g.H, g.W, S_, psi, hqs ? 1 : 0, x, fbp_, dqp_, dq_batch_, wid, eps, rho, rho_stride, it, tw_h_,
                  tw_w_, dps);
    } else {
      drv.xupdate(true, g.P - g.P % 2, 1, g.H, g.W, S_, psi, hqs ? 1 : 0, x, fbp_, dqp_, 1, wid, eps, rho, 0, it, tw_h_, tw_w_);
      if (g.P % 2) {
        float* x1 = x;
        const PsiPack one = last_plane(psi, &x1);
        drv.xupdate(false, 1, 1, g.H, g.W, S1_, one, hqs ? 1 : 0, x1, fbp1_, dqp1_, 1, wid, eps, rho, 0, it, tw_h_, tw_w_);
      }
    }
    return be.rc;
  }

 private:
  // tensor map of the pair spectrum S[pp][g][h][c] as {H*CG*2 floats (one column group), G, pairs}; box {ZR*CG*2 floats = 64 B, min(G,256), 1}
  int ensure_smap(int pairs, cudaStream_t s) {
    if (smap_dev_ && smap_pairs_ == pairs) return DPX_OK;
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return DPX_ERR_CUDA;
    }
    const int G = g_.W / fused::CG;
    const cuuint64_t dims[3] = {(cuuint64_t)g_.H * fused::CG * 2, (cuuint64_t)G, (cuuint64_t)pairs};
    const cuuint64_t strides[2] = {(cuuint64_t)g_.H * fused::CG * 8, (cuuint64_t)G * g_.H * fused::CG * 8};
    const cuuint32_t box[3] = {(cuuint32_t)(fused::ZR * fused::CG * 2), (cuuint32_t)fused::rowz_tma_box(G), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    alignas(64) CUtensorMap m;
    const CUresult r = reinterpret_cast<EncodeTiledFn>(fp)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, S_, dims, strides, box, estr,
                                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(S) failed with CUresult %d", (int)r); return DPX_ERR_CUDA; }
    if (!smap_dev_) DPX_CUDA(cudaMalloc(&smap_dev_, 128));
    DPX_CUDA(cudaMemcpyAsync(smap_dev_, &m, sizeof(m), cudaMemcpyHostToDevice, s));
    DPX_CUDA(cudaStreamSynchronize(s));               // `m` is a stack object (cold path: once per plan)
    smap_pairs_ = pairs;
    return DPX_OK;
  }
  static int upload_twiddles(int n, float2** out) {
    const std::vector<float2> t = twiddle_records_for(n);
    DPX_CUDA(cudaMalloc(out, t.size() * sizeof(float2)));
    DPX_CUDA(cudaMemcpy(*out, t.data(), t.size() * sizeof(float2), cudaMemcpyHostToDevice));
    return DPX_OK;
  }
  Geom g_{};
  FftEngine* inner_ = nullptr;
  float2 *S_ = nullptr, *fbp_ = nullptr, *tw_h_ = nullptr, *tw_w_ = nullptr;
  float* dqp_ = nullptr;
  size_t dqp_cap_ = 0, bytes_ = 0;
  int dq_batch_ = 1;
  int persist_ctas_ = 0;
  int col_tma_sms_ = 0;
  int* row_ctr_ = nullptr;
  int row_dyn_ = 1, col_bulk_ = 0, row_tma_ = 0, pdl_ = 0;
  void* smap_dev_ = nullptr;
  int smap_pairs_ = 0;
  static constexpr size_t kTraceRecs = 16384;          // >= CTAs of k_col / tiles of the row kernel at the traced batch
  const char* trace_path_ = nullptr;
  unsigned long long *trace_col_ = nullptr, *trace_row_ = nullptr;
  const float2* fb_std_ = nullptr;
  const float* dq_std_ = nullptr;
  const float* dpsi_std_ = nullptr;
  float* dpsp_ = nullptr;
  bool dpsi_dirty_ = false;
  bool fb_dirty_ = true, dq_dirty_ = true, dq_set_ = false;
  bool pairs_enabled_ = true, channel_shared_ = false;
  Mode packed_mode_ = PLANES;
  float2 *S1_ = nullptr, *fbp1_ = nullptr;
  float* dqp1_ = nullptr;
  cudaStream_t side_ = nullptr;
  cudaEvent_t ev_fork_ = nullptr, ev_join_ = nullptr;
};

}  // namespace

int make_fused_engine(const Geom& g, FftEngine** out) {
  *out = nullptr;
  if (!fused::size_supported_h(g.H) || !fused::size_supported(g.W)) {
    set_error("fused FFT engine: shape [%d x %d] not supported (either side 2^k in 64..4096, 3*2^k in 192..3072, 5*2^k in 320..2560, 960, 1600, "
              "1920, 3840; height also 720, 1080, 1200, 1440, 2160)", g.H, g.W);
    return DPX_ERR_INVALID;
  }
  FusedEngine* e = new (std::nothrow) FusedEngine();
  if (!e) { set_error("out of host memory"); return DPX_ERR_NOMEM; }
  int rc = e->init(g);
  if (rc) { e->destroy(); return rc; }
  *out = e;
  return DPX_OK;
}

}  // namespace dpx
