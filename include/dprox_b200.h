/*
 * dprox_b200.h — C-ABI of the B200-native proximal-iteration backend (libdprox_b200.so).
 *
 * Drop-in boundary for the per-iteration hot loop of Delta-Prox's compiled ADMM / LADMM / HQS /
 * ADMM_vxu / PGD solvers.  The reference has NO native boundary (it is eager PyTorch); each entry
 * point below names the reference Python code whose arithmetic it replaces (paths relative to
 * /root/reference).  INTEGRATION.md shows the ctypes binding a maintainer adds under dprox/algo.
 *
 * Conventions
 *   - Plain C: pointers + sizes only, no torch / C++ types.  All tensor pointers are DEVICE
 *     pointers to contiguous fp32 [B,C,H,W] arrays owned by the caller (PyTorch), unless a
 *     parameter is explicitly documented as a HOST pointer (`*_host`).
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *     Every call only enqueues work on that stream; nothing synchronises unless documented.
 *   - Return value: 0 = DPX_OK, otherwise a dpx_status; dpx_last_error() gives a thread-local
 *     message.  Nothing throws across the ABI.
 *   - A plan is bound to the device current at creation and may be driven by one host thread at
 *     a time.  The library owns plan-internal scratch (spectra, FFT plans, constants) only.
 *   - Spectra use the R2C half-spectrum layout [B,C,H,W/2+1] complex64 (interleaved re,im).
 */
#ifndef DPROX_B200_H_
#define DPROX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPX_ABI_VERSION 1
#define DPX_MAX_PSI 8

typedef enum {
  DPX_OK = 0,
  DPX_ERR_INVALID = 1,      /* bad argument / unsupported combination */
  DPX_ERR_CUDA = 2,         /* a CUDA runtime call failed              */
  DPX_ERR_CUFFT = 3,        /* a cuFFT call failed                     */
  DPX_ERR_STATE = 4,        /* constants not set / wrong call order    */
  DPX_ERR_NOMEM = 5
} dpx_status;

/* algo/{admm,hqs,pgd}.py — which `_iter` the plan executes */
typedef enum {
  DPX_ALGO_ADMM = 0,        /* ADMM._iter            algo/admm.py:49-59   */
  DPX_ALGO_HQS = 1,         /* HQS._iter             algo/hqs.py:10-16    */
  DPX_ALGO_ADMM_VXU = 2,    /* ADMM_vxu._iter        algo/admm.py:107-120 */
  DPX_ALGO_PGD = 3,         /* PGD._iter             algo/pgd.py:39-43    */
  DPX_ALGO_LADMM = 4        /* LinearizedADMM._iter  algo/admm.py:79-100 (identity psi linops only) */
} dpx_algo;

/* proxfn/sum_square.py:115-156 — which closed form the x-update uses */
typedef enum {
  DPX_X_FREQ_DIAG = 0,      /* solve_direct, freq branch   sum_square.py:150-152 */
  DPX_X_SPATIAL_DIAG = 1    /* solve_direct, spatial branch sum_square.py:154    */
} dpx_xupdate;

/* proxfn/{nonneg,norm}.py `_prox` bodies (+ box / external hooks) */
typedef enum {
  DPX_PROX_NONNEG = 0,      /* max(v,0)                    proxfn/nonneg.py:10-11 */
  DPX_PROX_L1 = 1,          /* sign(v) max(|v|-lam,0)      proxfn/norm.py:6-19    */
  DPX_PROX_L2SQ = 2,        /* v/(1+2 lam)                 proxfn/norm.py:22-27   */
  DPX_PROX_BOX = 3,         /* clamp(v, lo, hi)            (new; north_star)      */
  DPX_PROX_EXTERNAL = 4,    /* caller evaluates it (deep_prior: proxfn/pnp/prior.py:73-86) */
  DPX_PROX_ISO_TV = 5       /* isotropic TV: group shrink max(1 - lam/|w|_2, 0) w over the (grad_H, grad_W) pair;
                             * only with DPX_LINOP_GRAD_HW (new; north_star — the reference has no isotropic TV)  */
} dpx_prox_kind;

/* linop of a psi term, applied to the single variable x */
typedef enum {
  DPX_LINOP_IDENTITY = 0,   /* Variable / scale            linop/variable.py, linop/scale.py */
  DPX_LINOP_GRAD_H = 1,     /* grad(x, dim=0): x[i+1]-x[i] along H, circular  linop/grad.py:8-23 */
  DPX_LINOP_GRAD_W = 2,     /* grad(x, dim=1): along W                                             */
  DPX_LINOP_GRAD_HW = 3     /* [grad_H x ; grad_W x] stacked on the channel axis: state v,u are [B,2C,H,W] */
} dpx_linop_kind;

typedef struct {
  int32_t prox;             /* dpx_prox_kind */
  int32_t linop;            /* dpx_linop_kind */
  float scale;              /* scalar multiplying the linop (c * linop -> scale node)  */
  float alpha;              /* ProxFn.alpha   proxfn/base.py:78-82                      */
  float beta;               /* ProxFn.beta    proxfn/base.py:18-21                      */
  float box_lo, box_hi;     /* DPX_PROX_BOX only */
} dpx_psi_desc;

typedef struct {
  int32_t abi_version;      /* = DPX_ABI_VERSION */
  int32_t batch, channels, height, width;
  int32_t algo;             /* dpx_algo    */
  int32_t xupdate;          /* dpx_xupdate */
  int32_t n_psi;            /* 0..DPX_MAX_PSI */
  dpx_psi_desc psi[DPX_MAX_PSI];
  float eps;                /* 1e-7 in the reference (sum_square.py:115)            */
  int32_t fft_backend;      /* 0 = auto, 1 = cuFFT, 2 = fused sm_100a FFT kernels   */
  int32_t eps_delta;        /* SPATIAL_DIAG only: add eps to element [0,0] of every plane's numerator.  This is
                             * what the reference's Fourier branch computes when every diagonal is constant
                             * (all linops are scaled identities): F^-1[(F(n)+eps)/(d+eps)] = (n + eps*delta)/(d+eps),
                             * so such problems need no FFT at all (sum_square.py:150-152, tests/problem/
                             * test_ml_problems.py:5-22 expect x == rhs/2 exactly).                      */
} dpx_problem_desc;

typedef struct dpx_plan dpx_plan;

/* ---- library ------------------------------------------------------------------------------- */
int dpx_abi_version(void);
const char* dpx_last_error(void);
const char* dpx_build_info(void);          /* "sm_100a nvcc 12.9 ..." */
unsigned long long dpx_launch_count(void); /* library kernels launched so far by this process */

/* ---- plan lifetime ------------------------------------------------------------------------ */
/* Replaces compile()-time analysis of least_squares.__init__ (sum_square.py:87-110). */
int dpx_plan_create(const dpx_problem_desc* desc, dpx_plan** out);
void dpx_plan_destroy(dpx_plan* plan);
/* bytes of device scratch the plan holds */
size_t dpx_plan_workspace_bytes(const dpx_plan* plan);

/* ---- iteration-invariant constants (hoisted out of solve_direct, sum_square.py:125-148) ---- */
/* FREQ_DIAG.  ktb      : real [B,C,H,W]  = sum_q A_q^T b_q  (may be NULL = 0)
 *             dq       : real [dq_batch,C,H,W/2+1] = sum_q |OTF_q|^2  (dq_batch = 1 or B; NULL = 0)
 *             dpsi     : real [1,C,H,W/2+1] = sum over NON-identity psi terms of scale^2 |OTF_i|^2 or
 *                        NULL; identity psi terms contribute sum scale_i^2 analytically.
 * The library transforms ktb once (R2C) and keeps F(ktb); the arrays are copied. */
int dpx_plan_set_freq_constants(dpx_plan* plan, const float* ktb, const float* dq, int dq_batch,
                                const float* dpsi, void* stream);
/* Replace only the right-hand side K^T b (a new batch of measurements through the same operators): re-transforms
 * ktb, keeps the diagonals.  ktb real [B,C,H,W] or NULL (= 0).  Valid for both x-update kinds. */
int dpx_plan_set_rhs(dpx_plan* plan, const float* ktb, void* stream);
/* Hints about the constants the caller knows and the library would have to synchronise to find out.
 * DPX_HINT_CHANNEL_SHARED_DIAG = 1: the quadratic diagonal dq is identical for every channel (a single-channel PSF
 * broadcast over RGB, psf2otf.py:29): the fused engine may then pair ANY two planes, e.g. the channels of a single image. */
#define DPX_HINT_CHANNEL_SHARED_DIAG 1
int dpx_plan_set_hint(dpx_plan* plan, int hint, int value);
/* Same for the common data term sum_squares(scale * conv(x) - b): F(K^T b) = scale * conj(OTF) * F(b) is formed directly
 * in the Fourier domain (one R2C + one product) instead of conv.adjoint(b) (R2C, product, C2R) followed by another R2C.
 * b: real [B,C,H,W]; otf: complex64 half spectrum [otf_batch,C,H,W/2+1].  (sum_square.py:127-132, conv.py:37-41) */
int dpx_plan_set_rhs_spectral(dpx_plan* plan, const float* b, const float* otf, int otf_batch, float scale,
                              void* stream);
/* SPATIAL_DIAG. ktb real [B,C,H,W]; dq real [dq_batch,C,H,W] (mask diagonal). */
int dpx_plan_set_spatial_constants(dpx_plan* plan, const float* ktb, const float* dq, int dq_batch,
                                   void* stream);
/* SPATIAL_DIAG, optional: dpsi real [C,H,W] = sum over psi terms of scale_i^2 * diag_i (mask-type psi linops such as
 * norm1(mosaic(x)), or every psi term when the plan has n_psi = 0 and is driven through dpx_xsolve); the x-update divides by
 * dq + rho (dpsi + wid) + eps (sum_square.py:142-148, 154).  NULL clears. */
int dpx_plan_set_spatial_psi_diag(dpx_plan* plan, const float* dpsi, void* stream);
/* Which transform engine the plan's last fused call ran on: introspection for tests and bench lines. */
#define DPX_ENGINE_NONE (-1)        /* SPATIAL_DIAG plan: no transform */
#define DPX_ENGINE_CUFFT 0          /* cuFFT R2C / C2R + element-wise kernels (any size) */
#define DPX_ENGINE_FUSED_PLANES 1   /* fused sm_100a FFT kernels, half-spectrum planes */
#define DPX_ENGINE_FUSED_PAIRS 2    /* fused kernels, two planes per complex transform (even batch) */
#define DPX_ENGINE_FUSED_FLAT 3     /* fused kernels, planes paired across channels (+ odd last plane on the half-spectrum engine) */
int dpx_plan_engine_mode(const dpx_plan* plan);
/* Optional constant inside psi term i's linop (`norm1(x - c)`), real [B,C,H,W]; NULL clears. */
int dpx_plan_set_psi_offset(dpx_plan* plan, int i, const float* c, void* stream);

/* ---- the hot loop ------------------------------------------------------------------------- */
/* State layout per algorithm (all [B,C,H,W] fp32, updated in place):
 *   ADMM / LADMM : x, v[n_psi], u[n_psi]        HQS : x, v[n_psi] (u = NULL)
 *   ADMM_VXU     : x(=z), v[n_psi](=x_i), u[n_psi]          PGD : x only
 * rho  : device [T] (rho_stride = 0) or [B,T] row-major (rho_stride = T) schedule.
 * lam  : host array of n_psi device pointers, each [T] or [B,T] (lam_stride[i] = 0 or T).
 * Runs iterations it0 .. it0+n_iters-1 of the schedule (Algorithm.iters, algo/base.py:128-156).
 * resid: optional device [n_iters,B,4] = {|r|^2, |s|^2, |Kx|^2, |v|^2} per sample (NULL = skip).
 * Plans containing a DPX_PROX_EXTERNAL term must be driven stage-wise (dpx_stage_*). */
int dpx_iters(dpx_plan* plan, float* x, float* const* v, float* const* u,
              const float* rho, int rho_stride, const float* const* lam, const int* lam_stride,
              int it0, int n_iters, float* resid, void* stream);

/* Stage-wise form of one iteration, for external prox terms / Python callbacks:
 *   dpx_stage_xupdate : x <- least_square.solve(b, rho)     (sum_square.py:115-156)
 *   dpx_stage_prox    : v_i <- prox_i(K_i x + u_i, lam_i), u_i <- u_i + K_i x - v_i for NATIVE terms,
 *                       and w_i = K_i x + u_i written to v_i for EXTERNAL terms (the caller then
 *                       evaluates its prox on w_i and calls dpx_stage_dual_external).
 *   PGD: dpx_stage_xupdate writes the gradient step x - rho*grad f(x) into v[0] (caller scratch),
 *        dpx_stage_prox writes x <- prox(v[0], lam). */
int dpx_stage_xupdate(dpx_plan* plan, float* x, float* const* v, float* const* u,
                      const float* rho, int rho_stride, int it, void* stream);
int dpx_stage_prox(dpx_plan* plan, float* x, float* const* v, float* const* u,
                   const float* const* lam, const int* lam_stride, int it, void* stream);
/* x <- closed-form least_squares.solve given the caller-computed psi part of the right-hand side
 * t = sum_i A_i^T b_i (real [B,C,H,W]):  FREQ: F^-1[(F(ktb) + rho F(t) + eps)/(dq + rho(dpsi+wid) + eps)],
 * SPATIAL: (ktb + rho t)/(dq + rho wid + eps).  Lets the host compose LADMM / ADMM_vxu / external-prox
 * variants from the stand-alone kernels below.  (sum_square.py:123-156) */
int dpx_xsolve(dpx_plan* plan, const float* t, const float* rho, int rho_stride, int it, float* x, void* stream);
/* Backward of dpx_xsolve, the closed form of what the reference obtains by autograd through
 * least_squares.solve_direct (sum_square.py:123-156; used by unrolled training, specialization/unroll.py:42-58):
 * given g = dL/dx and the forward output x,
 *   g_ktb = F^-1[F(g) / Dn]                 (= dL/d(sum_q A_q^T b_q);  dL/dt = rho * g_ktb)
 *   g_rho = sum_k Re(conj(F g)_k (F(t) - (dpsi+wid) F(x))_k / Dn_k) / (H W)      [B] if rho_stride != 0, else [1]
 * with Dn = dq + rho (dpsi + wid) + eps.  g_rho may be NULL.  x is only read when g_rho is requested.
 * SPATIAL_DIAG plans: g_ktb = g / D, g_rho = sum g_ktb (x (dq + eps) - ktb) / rho with the same D in pixel space. */
int dpx_xsolve_backward(dpx_plan* plan, const float* g, const float* x, const float* rho, int rho_stride, int it,
                        float* g_ktb, float* g_rho, void* stream);
/* v_i <- K_i x0 (affine: scale * A_i x0 - c_i), u_i <- 0.   ADMM.initialize / HQS.initialize
 * (algo/admm.py:61-67, algo/hqs.py:5-8).  u may be NULL for HQS. */
int dpx_init_state(dpx_plan* plan, const float* x, float* const* v, float* const* u, void* stream);
/* u_i <- w_i - v_new (w_i currently stored in u_i's slot is NOT assumed): u_i <- u_i + Kx_i - v_new
 * given w = Kx_i + u_i:  u_i <- w - v_new;  v_i <- v_new.  (admm.py:56-57) */
int dpx_stage_dual_external(dpx_plan* plan, int i, const float* w, const float* v_new, float* v_i,
                            float* u_i, void* stream);

/* ---- stand-alone operator kernels (LinOp.forward/adjoint, ProxFn.prox outside the fused loop) */
/* y = Re F^-1( OTF (or conj OTF) * F x ) — conv.forward/adjoint, linop/conv.py:31-41.
 * otf: complex64 half spectrum [otf_batch,C,H,W/2+1], otf_batch in {1,B}. */
int dpx_spectral_filter(dpx_plan* plan, const float* x, const float* otf, int otf_batch, int conjugate,
                        float* y, void* stream);
/* out = prox(v, lam) with the ProxFn.prox wrapper chain (proxfn/base.py:55-64).
 * lam: device [B] (lam_per_sample=1) or [1].  offset may be NULL. */
int dpx_prox_apply(int prox_kind, const float* v, const float* lam, int lam_per_sample, float alpha,
                   float beta, float box_lo, float box_hi, const float* offset, float* out, int batch,
                   size_t per_sample, void* stream);
/* Backward of dpx_prox_apply for the native `_prox` bodies (nonneg / l1 / l2sq / box):
 * g_v = dprox/dv * g;  g_lam (device [B] or [1], may be NULL) = sum dprox/dlam * g. */
int dpx_prox_backward(int prox_kind, const float* v, const float* lam, int lam_per_sample, float alpha, float beta,
                      float box_lo, float box_hi, const float* offset, const float* g, float* g_v, float* g_lam,
                      int batch, size_t per_sample, void* stream);
/* out = a*x + b*y (+ c*z); coefficient pointers are device [B] or [1] arrays or NULL (=1). y,z may be NULL. */
int dpx_lincomb(float* out, const float* a, const float* x, const float* b, const float* y,
                const float* c, const float* z, int coeff_per_sample, int batch, size_t per_sample,
                void* stream);
/* out = a*x + b*y with HOST scalars (y may be NULL): scale.forward, sum.forward, `x - c` glue
 * (linop/scale.py:21-33, linop/sum.py:13-20). n = total element count. */
int dpx_axpby(float* out, float a, const float* x, float b, const float* y, size_t n, void* stream);
/* circular forward difference / its adjoint along H (axis=0) or W (axis=1): grad.py:8-23 */
int dpx_grad_apply(const float* x, float* y, int planes, int height, int width, int axis, int adjoint,
                   float scale, void* stream);

/* out[p][y][x] = in[p][y - top][x - left] where that lies inside the [h_in, w_in] input, else 0: the zero padding (top, left >= 0)
 * and the crop (negative offsets) of the `circular=False` convolutions (linop/conv.py:100-121, contrib/optic/common.py:97-117). */
int dpx_pad2d(const float* in, float* out, int planes, int h_in, int w_in, int h_out, int w_out, int top, int left,
              void* stream);
/* the 8 flips / rotations of the x8 test-time augmentation, Augment.augment (proxfn/pnp/denoisers/composite.py:30-47);
 * modes 1, 3, 5, 7 transpose: out is [planes, width, height]. */
int dpx_augment(const float* in, float* out, int planes, int height, int width, int mode, void* stream);

/* out = w * x — mosaic / mul_elementwise forward = adjoint (linop/subsample.py:18-31, linop/mul.py:59-65).
 * w: real [w_batch, per_sample], w_batch in {1, batch}. */
int dpx_mul_apply(float* out, const float* x, const float* w, int w_batch, int batch, size_t per_sample,
                  void* stream);
/* out[b] = max_i |x[b,i]|  (pcg's inf-norm stop test, solver_cg.py:225-229); out is device [batch]. */
int dpx_absmax(const float* x, float* out, int batch, size_t per_sample, void* stream);

/* ---- fused (P)CG vector kernels (linalg/solve/solver_cg.py:56-233) ------------------------- */
/* dots[b] = <x_b, y_b> per sample (bdot, solver_cg.py:7-22); dots is device [B] (zeroed inside). */
int dpx_cg_dot(const float* x, const float* y, float* dots, int batch, size_t per_sample, void* stream);
/* alpha_b = gamma_b / pq_b ; x += alpha p ; r -= alpha q ; gamma_new_b = <r_b, r_b>   (cg :125-129) */
int dpx_cg_update(float* x, float* r, const float* p, const float* q, const float* gamma, const float* pq,
                  float* gamma_new, int batch, size_t per_sample, void* stream);
/* beta_b = gamma_new_b / gamma_old_b ; p = r + beta p    (cg :113-116) */
int dpx_cg_direction(float* p, const float* r, const float* gamma_new, const float* gamma_old, int batch,
                     size_t per_sample, void* stream);

/* Device-side stop test of cg / pcg (solver_cg.py:103-107, 225-229), no host round trip: *done (device int, zeroed by the
 * caller before the solve) becomes and stays 1 once val[b] <= tol[b] (strict != 0: <) for every b in the batch; while it is
 * set, pq[b] is overwritten with +inf, which makes the following dpx_cg_update a no-op (alpha = 0, x and r frozen).
 * tol: device [tol_n], tol_n in {1, batch}. */
int dpx_cg_gate(const float* val, const float* tol, int tol_n, int strict, float* pq, int* done, int batch, void* stream);

/* ---- native deep-denoiser (FFDNet-color) on tcgen05 tensor cores ---------------------------------------------
 * deep_prior -> FFDNetColorDenoiser -> FFDNet.forward (proxfn/pnp/prior.py:73-86, denoisers/wrapper.py:38-48,
 * models/network_ffdnet.py:44-68).  Hand-written sm_100a kernel (csrc/dpx_conv_tc.cuh): tcgen05.mma cta_group::2, TMEM
 * accumulators, TMA-staged activation rows reused for all nine taps, filter bank resident in shared memory.
 * Two precisions (dpx_ffdnet_set_precision): 0 = bf16 operands, fp32 accumulation (the fast denoiser, ~1e-2 relative to the
 * fp32 network); 1 = every operand as fp16 hi + 2^-11 lo' pieces, three MMAs per k-step into two TMEM accumulators
 * (fp32-class: ~1e-6 relative to the fp32 network -- the mode that meets the 1e-5 parity bar of the reference's fp32 path). */
typedef struct dpx_ffdnet dpx_ffdnet;
int dpx_ffdnet_available(void);                       /* 1: the tensor-core path is always built */
int dpx_ffdnet_create(int nb, int nc, dpx_ffdnet** out);          /* nb conv layers, nc = 96 channels */
void dpx_ffdnet_destroy(dpx_ffdnet* net);
int dpx_ffdnet_set_precision(dpx_ffdnet* net, int mode);          /* 0 = bf16 (default), 1 = fp16-pair split (fp32-class) */
/* layer 0 = head [nc,13,3,3], 1..nb-2 = body [nc,nc,3,3], nb-1 = tail [12,nc,3,3]; w/bias: device fp32, nn.Conv2d layout */
int dpx_ffdnet_set_layer(dpx_ffdnet* net, int layer, const float* w, const float* bias, int cout, int cin, void* stream);
/* y = FFDNet(x, sigma); x, y device fp32 [B,3,H,W]; sigma device [B] (sigma_per_sample=1) or [1] */
int dpx_ffdnet_forward(dpx_ffdnet* net, const float* x, const float* sigma, int sigma_per_sample, float* y, int B, int H,
                       int W, void* stream);
/* Same, keeping every layer's activation for dpx_ffdnet_backward (unrolled training with a frozen denoiser: what the
 * reference obtains by autograd through FFDNet.forward, e2e_optics_dprox.py:34-58). */
int dpx_ffdnet_forward_train(dpx_ffdnet* net, const float* x, const float* sigma, int sigma_per_sample, float* y, int B, int H,
                             int W, void* stream);
/* Data gradient of the last dpx_ffdnet_forward_train call: g_x [B,3,H,W] = dL/dx and g_sigma ([B] or [1], may be NULL) =
 * dL/dsigma given g_y = dL/dy.  Every layer's input gradient is again a 3x3 convolution (transposed, flipped filter) on the
 * same tensor-core kernel, with the ReLU mask of the saved activation applied in its epilogue.  Consumes the saved state. */
int dpx_ffdnet_backward(dpx_ffdnet* net, const float* g_y, float* g_x, float* g_sigma, int sigma_per_sample, int B, int H,
                        int W, void* stream);
/* dpx_ffdnet_backward plus the gradients w.r.t. every layer's parameters (training the denoiser's weights without leaving
 * native code): gw[l] device fp32 [cout,cin,3,3] (nn.Conv2d layout), gb[l] device fp32 [cout]; gb or entries of it may be NULL.
 * Weight gradient on the MN-major tcgen05 kernel (csrc/dpx_conv_wgrad.cuh).  bf16 precision. */
int dpx_ffdnet_backward_params(dpx_ffdnet* net, const float* g_y, float* g_x, float* g_sigma, int sigma_per_sample,
                               float* const* gw, float* const* gb, int B, int H, int W, void* stream);
/* One convolution layer on fp32 NCHW tensors (per-layer parity tests): direction 0 = forward (+bias, optional ReLU),
 * 1 = data gradient.  x [B,cin,H,W] -> y [B,cout,H,W], (cin, cout) = the layer's channel counts in that direction. */
int dpx_ffdnet_conv_layer(dpx_ffdnet* net, int layer, int direction, int relu, const float* x, float* y, int B, int H, int W,
                          void* stream);
/* Weight and bias gradient of one layer on fp32 NCHW tensors (torch: conv2d's grad_weight / grad_bias; network_ffdnet.py:27-68
 * under training): x [B,cin,H,W] = the layer's input, gy [B,cout,H,W] = gradient w.r.t. its pre-activation output;
 * gw [cout,cin,3,3], gb [cout] (may be NULL).  tcgen05 kernel with MN-major operands (csrc/dpx_conv_wgrad.cuh), bf16 operands,
 * fp32 accumulation in TMEM over all pixels. */
int dpx_ffdnet_wgrad_layer(dpx_ffdnet* net, int layer, const float* x, const float* gy, float* gw, float* gb, int B, int H, int W,
                           void* stream);

/* ---- CS-MRI closed-form data term on complex iterates (proxfn/fast/csmri.py:14-25; the ext_sum_squares hook,
 * proxfn/sum_square.py:44-48).  v, y, out: complex64 [B,C,H,W] (interleaved re,im); mask: fp32 0/1 [mask_batch,C,H,W],
 * mask_batch in {1,B}; y and mask in the reference's CENTRED k-space convention (utils/misc.py:164-193).
 *   out = ifft2c( mask ? (rho fft2c(v) + y) / (1 + rho num_psi) : fft2c(v) ),  rho: device [B] or [1]. */
int dpx_csmri_prox(const float* v, const float* y, const float* mask, int mask_batch, const float* rho,
                   int rho_per_sample, float num_psi, float* out, int batch, int channels, int height, int width,
                   void* stream);
/* real [n] -> complex64 [n] (imaginary part 0) and the real part of complex64 [n]: glue for operators whose iterates
 * are complex while a prox (deep denoiser) works on the real part (pnp/prior.py:79). */
int dpx_real_to_complex(const float* x, float* out, size_t n, void* stream);
int dpx_complex_real(const float* z, float* out, size_t n, void* stream);

/* ---- DOE optics forward model feeding the unrolled solver (contrib/optic/{common,doe_model}.py) --------
 * Building blocks with their backward counterparts (the reference differentiates the pipeline by autograd).  Complex
 * arrays are complex64 (interleaved re,im). */
/* batched 2-D C2C transform of `planes` [H,W] arrays, unnormalised in both directions (torch.fft.fft2 / ifft2 * H*W);
 * its adjoint is the opposite direction.  FresnelPropagator.forward common.py:155-164, img_psf_conv common.py:107-110. */
int dpx_c2c(const float* in, float* out, int planes, int height, int width, int inverse, void* stream);
/* out[i] = scale * a[i] * b[i mod n_b] (b conjugated if conj_b): transfer-function / OTF products with a broadcast factor;
 * dpx_cmul_reduce gives the gradient of that broadcast factor, out[j] = scale * sum_k g[k n_b + j] conj(a[k n_b + j]). */
int dpx_cmul(const float* a, const float* b, float* out, size_t n_a, size_t n_b, int conj_b, float scale, void* stream);
int dpx_cmul_reduce(const float* g, const float* a, float* out, size_t n_b, int batch, float scale, void* stream);
/* field[l] = aperture * exp(i coef[l] h^2) on the N x N wavefront, zero-padded by `pad` on every side
 * (HeightMap.get_phase_profile doe_model.py:37-51 with coef[l] = k_l * (n_l - 1); aperture and padding of
 * RGBCollimator.get_psf :103-104 / FresnelPropagator.forward common.py:156-157).  h: [N,N], field: [L, N+2pad, N+2pad]. */
int dpx_doe_field(const float* h_sqrt, const float* coef, const float* aperture, float* field, int n_lambda, int n, int pad,
                  void* stream);
int dpx_doe_field_backward(const float* h_sqrt, const float* coef, const float* aperture, const float* g_field, float* g_h,
                           int n_lambda, int n, int pad, void* stream);
/* out[l] = scale * avg_pool_factor(|crop_pad(field[l])|^2): intensity + area_downsampling (doe_model.py:106-107,
 * common.py:27-44).  field: [L, N+2pad, N+2pad] complex, out: [L, N/factor, N/factor]. */
int dpx_abs2_pool(const float* field, float* out, int n_lambda, int n, int pad, int factor, float scale, void* stream);
int dpx_abs2_pool_backward(const float* field, const float* g, float* g_field, int n_lambda, int n, int pad, int factor,
                           float scale, void* stream);
/* out = x / sum(x) (psfs / psfs.sum(), doe_model.py:109); sum_out: device scalar kept for the backward
 * g_x = (g - <g, out>) / sum  (scratch1: device scalar). */
int dpx_normalize_sum(const float* x, float* out, float* sum_out, size_t n, void* stream);
int dpx_normalize_sum_backward(const float* g, const float* out, const float* sum, float* scratch1, float* g_x, size_t n,
                               void* stream);

/* ---- end-to-end host-buffer entry point (the e2e leg of bench.py) -------------------------- */
/* Copies x0 (HOST, pinned or pageable, [B,C,H,W]) to the device, initialises (v = K x0, u = 0),
 * runs n_iters iterations with HOST schedules rho_host [T] / lam_host [n_psi][T] (scalars per
 * iteration), copies x back to x_out_host and synchronises the stream.  (Algorithm.solve, base.py:85-126) */
int dpx_solve_host(dpx_plan* plan, const float* x0_host, float* x_out_host, const float* rho_host,
                   const float* lam_host, int n_iters, void* stream);

/* ---- residual-based stop criterion support (new, opt-in; SURVEY App. C "Residuals") ------- */
/* Reduces resid [n,B,4] rows into per-iteration totals on device: out[n,4] = sum_b resid[n,b,:]. */
int dpx_resid_reduce(const float* resid, float* out, int n, int batch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPROX_B200_H_ */
