// dpx_conv_tc.cuh — hand-written 3x3 / stride 1 / pad 1 convolution on the 5th-generation tensor cores (sm_100a):
// the FFDNet layers behind `deep_prior` (proxfn/pnp/denoisers/models/network_ffdnet.py:27-68), forward and data gradient.
//
// Formulation.  Implicit GEMM  D[pixel, cout] = sum_{tap, cin} A_tap[pixel, cin] * W_tap[cout, cin]  with
//   * activations in HBM as channel-group-major bf16 with one explicit zero column on either side of every row,
//       X[n][cg][h][row_pitch(W)][8]   (cg = channel / 8: 16 bytes per pixel and group; image pixel x lives at index x + 1),
//   * one MMA = 2 x 128 pixels of one image row each, issued for a CTA PAIR (tcgen05 cta_group::2, M = 256, N = cout): each
//     CTA owns 128 pixels (its TMEM lanes) and HALF of the filter bank (N/2 output channels), so the whole 3x3xCinxCout
//     filter (162 KB for 96 -> 96) stays RESIDENT in shared memory for the lifetime of the kernel,
//   * the activation rows a tile needs (y-1, y, y+1; 130 pixels each = 128 + halo) live in a ring of row slots filled by
//     TMA (cp.async.bulk.tensor.5d; rows above / below the image = out-of-bounds zero fill, the left / right padding = the
//     zero columns of the layout) and are REUSED for all nine taps and by three consecutive output rows: a tap is nothing
//     but a different start address of the same shared-memory tile in the MMA's matrix descriptor (K-major, no swizzle:
//     pixel stride 16 B, channel-group stride 130 * 16 B), so every activation is fetched from L2 once per row block
//     (+ 2/16 halo) instead of once per tap.  (A TMA box dimension is limited to 256 elements and short inner runs are an
//     order of magnitude slower, so the 130-pixel run of one channel group is described as 2 x 65 pixels of 8-byte
//     elements with a separate unit-stride "start pixel" dimension.)
//   * fp32 accumulators in TMEM, three stages (the MMAs of rows y+1, y+2 overlap the epilogue of row y),
//   * epilogue (8 warps: two per TMEM lane quadrant, 48 channels each): tcgen05.ld x3 -> + bias -> ReLU (or the ReLU mask of
//     the saved forward activation: data gradient) -> bf16 -> coalesced 16-byte stores in the same layout (the next
//     layer's input).
// SPLIT (fp32-class accuracy on the same tensor cores, `precision="fp32"` of the denoiser): every operand is the sum of two fp16
// numbers, v = hi + 2^-11 lo' (hi = fp16(v), lo' = fp16((v - hi) 2^11): 22 mantissa bits), and the product is formed as
//     main = a_hi w_hi            corr = a_lo' w_hi + a_hi w_lo'            result = main + 2^-11 corr
// in TWO TMEM accumulators per stage (the lo' lo' term is 2^-22 relative and dropped; fp16 x fp16 products are exact in the
// fp32 accumulator).  Three MMAs per k-step instead of one.  Both operand pieces of a layer do not fit shared memory next to
// each other for 96 -> 96 channels, so such a layer runs as TWO launches over one half of the input channels each (ring slot =
// hi + lo' pieces of 48 channels = the same 12 planes as the bf16 kernel; filter image = hi + lo' pieces of that half = the same
// 83 KB), handing an fp32 partial sum from the first to the second launch, which finalises (bias, ReLU / mask, re-split).
// Activation tensors of this mode: fp16 [N][khalf][piece][kp][H][W+2][8] with kp = channel groups per K-half.
// Warp roles per CTA (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocation + (leader CTA) MMA issue,
// warps 2..9 = epilogue.  Persistent: a cluster walks over pairs of (image, 128-pixel column tile, 16-row block) work units.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpx {
namespace convtc {

constexpr int TILE_PX = 128;             // output pixels per CTA and tile row (= UMMA M per CTA)
// pixels per padded row of an activation tensor: one zero column on either side, and the row is extended with zeros to a whole
// number of 128-pixel tiles so that a staged 130-pixel run never reaches into the next row (the weight-gradient kernel SUMS over
// every staged pixel; the forward kernel merely discarded what it computed from the overhang)
__host__ __device__ constexpr int row_pitch(int W) { return (W + TILE_PX - 1) / TILE_PX * TILE_PX + 2; }
constexpr int HALO_PX = TILE_PX + 2;     // staged pixels per row
constexpr int ROW_BLOCK = 16;            // output rows per work unit
constexpr int NSLOT = 5;                 // activation row slots in the ring (3 live + 2 in flight)
constexpr int N_EPI_WARPS = 8;            // bf16 kernel; the SPLIT kernel of a 96-channel layer runs 12 (Cfg::NTHREADS)
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> the even (leader) CTA

template <int CGIN, int COUT, bool SPLIT = false>
struct Cfg {
  static constexpr int NH = COUT / 2;                                   // output channels held by one CTA of the pair
  static constexpr int KSTEPS = CGIN / 2;                               // UMMA_K = 16 bf16 = 2 channel groups
  static constexpr uint32_t ROW_BYTES = CGIN * HALO_PX * 16;            // one staged activation row
  static constexpr uint32_t SLOT_BYTES = (ROW_BYTES + 127) / 128 * 128;  // ring-slot stride (TMA destinations are 128-byte aligned)
  static constexpr uint32_t W_TAP_BYTES = CGIN * NH * 16;               // one tap of this CTA's filter half
  static constexpr uint32_t W_BYTES = 9 * W_TAP_BYTES;
  static constexpr uint32_t A_LBO = HALO_PX * 16, B_LBO = NH * 16;      // byte distance of the two K halves of one MMA
  static constexpr int CORR_OFF = COUT <= 32 ? 32 : 128;                // SPLIT: TMEM column offset of the correction accumulator
  static constexpr int ACC_STRIDE = (SPLIT ? 2 : 1) * CORR_OFF;         // TMEM columns per accumulator stage
  static constexpr int NACC = (SPLIT && COUT > 32) ? 2 : 3;             // accumulator stages
  static constexpr int TMEM_COLS = COUT <= 32 ? (SPLIT ? 256 : 128) : 512;   // power of two >= NACC * ACC_STRIDE
  static constexpr int KP = CGIN / 2;                                   // SPLIT: channel groups per operand piece
  static constexpr int OUT_KP = COUT >= 96 ? 6 : COUT / 8;              // SPLIT: groups per piece and K-half of the OUTPUT tensor (a
                                                                        // 96-channel tensor is consumed in two K-halves of 6 groups)
  static_assert(!SPLIT || CGIN % 4 == 0, "SPLIT: hi and lo' pieces of an even number of channel groups");
  // epilogue warps per TMEM lane quadrant: 2 (1 when the layer is too narrow to split); the SPLIT epilogue of a 96-channel layer
  // (two accumulators, the partial sum of the other K-half, the re-split) was the bottleneck of its launch with 2 -> 3 x 32 channels
  static constexpr int EPI_SPLIT = (COUT == 96) ? 3 : (COUT / 2 >= 16 ? 2 : 1);
  static constexpr int EPI_COLS = COUT / EPI_SPLIT;                     // channels per epilogue warp
  static constexpr int NTHREADS = 64 + 32 * 4 * (EPI_SPLIT > 2 ? EPI_SPLIT : 2);
  static_assert(EPI_COLS % 16 == 0, "tcgen05.ld x16");
  static constexpr size_t SMEM = 1024 + (W_BYTES + 127) / 128 * 128 + (size_t)NSLOT * SLOT_BYTES + 256;
  static_assert(CGIN % 2 == 0 && COUT % 16 == 0 && COUT <= 128, "shape");
  static_assert(W_TAP_BYTES % 16 == 0, "alignment");
};

struct Params {
  const void* wpack;               // [2 halves][9 taps][CGIN][NH][8] bf16 (SPLIT: fp16, groups = hi piece then lo' piece)
  float bias[96];                  // per output channel (kernel-parameter constant bank: free operands in the epilogue)
  __nv_bfloat16* out;              // [N][COUT/8][H][W+2][8]
  const __nv_bfloat16* mask;       // optional, same layout as out: result is zeroed where mask <= 0 (ReLU backward)
  int N, H, W;
  int relu;
  int n_tiles, x_tiles, row_blocks;   // CTA work units: (image, 128-pixel column tile, 16-row block)
  int in_planes, in_plane_off;     // planes per image of the input tensor; first plane of this launch (SPLIT: khalf * CGIN)
  // ---- SPLIT mode ----
  __half* out16;                   // output pieces [N][khalf][piece][out_kp][H][W+2][8] (next layer's input / saved activation)
  const __half* mask16;            // optional: saved forward activation in the out16 layout; zero the result where its hi piece <= 0
  int out_kp;                      // channel groups per piece and K-half of the output tensor (6 for 96 channels, 2 for 16)
  const float* pin;                // fp32 partial sum of the previous K-half [N][COUT/8][H][W+2][8], or nullptr
  float* pout;                     // if set: write the fp32 partial sum there and do not finalise
  float* out32;                    // if set (last layer): fp32 result [N][COUT/8][H][W+2][8] instead of pieces
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(b)), "r"(count)); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n"
      "D_%=:\n\t}" ::"r"(s2u(b)), "r"(parity) : "memory");
}
// arrive on the LEADER CTA's copy of a barrier (same offset in its shared memory).  RELAXED: the only thing the arrival publishes
// is "this warp's tcgen05.ld of the accumulator has completed", which tcgen05.wait::ld + tcgen05.fence::before_thread_sync already
// order; a release at cluster scope would also wait for the warp's outstanding global stores (MEMBAR.ALL.GPU: it was 55 % of all
// stall samples of the first version)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* b) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(s2u(b) & PEER_MASK) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s2u(dst)), "l"(src),
               "r"(bytes), "r"(s2u(bar)) : "memory");
}
// one staged activation row of this CTA: tensor dims {130 x 8 B = 65 px, 2 halves, start pixel, row, plane = image * CG + cg},
// box {130, 2, 1, 1, CGIN} at (0, 0, xs, y, plane0); completion bytes are counted on the LEADER's barrier (cta_group::2 form),
// which the single MMA-issuing thread of the pair waits on
__device__ __forceinline__ void tma_row(void* dst, const CUtensorMap* map, uint64_t* bar, int xs, int y, int plane0) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(s2u(dst)), "l"(map), "r"(s2u(bar) & PEER_MASK), "r"(0), "r"(0), "r"(xs), "r"(y),
      "r"(plane0)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, int cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(dst_smem)), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_free(uint32_t addr, int cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], M = 256 over the CTA pair, K = 16 bf16
__device__ __forceinline__ void umma_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
      "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrive on barrier `b` of BOTH CTAs once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit_both(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s2u(b)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
}
// (lo, hi) fp32 -> packed fp16, round to nearest even, saturating to the largest finite value
__device__ __forceinline__ __half2 cvt_sat_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle: rows (pixels / output channels) 16 B apart, 8-row groups `sbo` bytes
// apart, the two 8-element K halves `lbo` bytes apart (cute/arch/mma_sm100_desc.hpp: bits [0,14) addr >> 4, [16,30) LBO >> 4,
// [32,46) SBO >> 4, [46,48) version = 1, [61,64) layout = 0)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}
// instruction descriptor of kind::f16: fp32 accumulate, bf16 x bf16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, bool fp16 = false) {
  // a_format [7,10) / b_format [10,13): 0 = fp16, 1 = bf16
  return (1u << 4) | (fp16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Unit { int n, x0, y0; };
// CTA `rank` of cluster-level work item `u` takes CTA tile 2u + rank; tiles beyond the end are dummies (n = N: every load is out
// of bounds = zeros, nothing is stored), which keeps the two CTAs of a pair in lockstep
__device__ __forceinline__ Unit unit_of(const Params& P, int u, int rank) {
  Unit t;
  const int tile = 2 * u + rank;
  const int xt = tile % P.x_tiles;
  const int rb = (tile / P.x_tiles) % P.row_blocks;
  t.n = tile / (P.x_tiles * P.row_blocks);
  t.x0 = xt * TILE_PX;
  t.y0 = rb * ROW_BLOCK;
  return t;
}

template <int CGIN, int COUT, bool SPLIT = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((Cfg<CGIN, COUT, SPLIT>::NTHREADS), 1)
    k_conv3x3_tc(const __grid_constant__ CUtensorMap in_map, Params P) {
  using C = Cfg<CGIN, COUT, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sm_w = base;                                       // this CTA's filter half: [9][CGIN][NH][8] bf16
  uint8_t* sm_a = base + (C::W_BYTES + 127) / 128 * 128;      // ring: [NSLOT][CGIN][130][8] bf16
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_a + (size_t)NSLOT * C::SLOT_BYTES);
  uint64_t* full = bars;                                      // [NSLOT]  (leader's copy is the live one)
  uint64_t* empty = bars + NSLOT;                             // [NSLOT]  per CTA
  uint64_t* tfull = bars + 2 * NSLOT;                         // [NACC]   per CTA
  uint64_t* tempty = bars + 2 * NSLOT + C::NACC;              // [NACC]   leader's copy
  uint64_t* wbar = bars + 2 * NSLOT + 2 * C::NACC;            // filter bank landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSLOT + 2 * C::NACC + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int n_units = (P.n_tiles + 1) / 2;
  constexpr int EPI_WARPS_USED = 4 * C::EPI_SPLIT;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < C::NACC; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 2 * EPI_WARPS_USED); }
    mbar_init(wbar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (warp == 0 && lane == 0) {                               // resident filter half of this CTA
    asm volatile("prefetch.tensormap [%0];" ::"l"(&in_map) : "memory");
    mbar_expect_tx(wbar, C::W_BYTES);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(P.wpack) + (size_t)rank * C::W_BYTES;
    for (int t = 0; t < 9; ++t) bulk_g2s(sm_w + t * C::W_TAP_BYTES, src + (size_t)t * C::W_TAP_BYTES, C::W_TAP_BYTES, wbar);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  mbar_wait(wbar, 0);
  tc_fence_before();
  cluster_sync_all();                                         // barriers initialised, both filter halves resident, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: one activation row per step into the next ring slot ===========================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int u = cluster_id; u < n_units; u += n_clusters) {
        const Unit t = unit_of(P, u, rank);
        for (int r = 0; r < ROW_BLOCK + 2; ++r, ++it) {
          const uint32_t s = it % NSLOT, ph = (it / NSLOT) & 1;
          mbar_wait(empty + s, ph ^ 1);
          if (leader) mbar_expect_tx(full + s, 2 * C::ROW_BYTES);
          // padded index of image pixel x0 - 1 is x0
          tma_row(sm_a + (size_t)s * C::SLOT_BYTES, &in_map, full + s, t.x0, t.y0 - 1 + r, t.n * P.in_planes + P.in_plane_off);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA): 9 taps x KSTEPS k-steps per output row.  The WHOLE warp runs the loop so that addresses and
    // descriptors stay in uniform registers; only the tcgen05 instructions themselves are predicated on one elected lane (a
    // single-lane branch made every operand a per-thread value and cost a ~20-instruction uniform-broadcast loop per MMA:
    // 91 cycles per MMA instead of the tensor pipe's 48) =====
    if (leader) {
      constexpr uint32_t IDESC = instr_desc(2 * TILE_PX, COUT, SPLIT);
      const uint32_t a0 = s2u(sm_a);
      const uint64_t bd0 = smem_desc(s2u(sm_w), C::B_LBO, 128);
      uint32_t it_base = 0, acc_it = 0;
      for (int u = cluster_id; u < n_units; u += n_clusters) {
        int waited = 0;
        for (int j = 0; j < ROW_BLOCK; ++j, ++acc_it) {
          const uint32_t a = acc_it % C::NACC, aph = (acc_it / C::NACC) & 1;
          mbar_wait(tempty + a, aph ^ 1);
          while (waited < j + 3) {
            const uint32_t i = it_base + waited;
            mbar_wait(full + (i % NSLOT), (i / NSLOT) & 1);
            ++waited;
          }
          tc_fence_after();
          const uint32_t d = tmem_base + a * C::ACC_STRIDE;
          // descriptors differ from a per-slot / per-bank base only in the 14-bit start-address field (bytes >> 4; shared
          // memory addresses stay below 2^18, so the field never carries): one add per operand and MMA
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint32_t slot = (it_base + j + dy) % NSLOT;
            const uint64_t ad0 = smem_desc(a0 + slot * C::SLOT_BYTES, C::A_LBO, 128);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              if constexpr (!SPLIT) {
#pragma unroll
                for (int k = 0; k < C::KSTEPS; ++k) {
                  const uint64_t ad = ad0 + (uint64_t)((dx * 16 + 2 * k * C::A_LBO) >> 4);
                  const uint64_t bd = bd0 + (uint64_t)(((dy * 3 + dx) * C::W_TAP_BYTES + 2 * k * C::B_LBO) >> 4);
                  if (elect_one()) umma_2sm(d, ad, bd, IDESC, (dy | dx | k) != 0);
                }
              } else {
                // groups [0, KP) = hi piece, [KP, 2 KP) = lo' piece, of the activations (ring slot) and of the filter image alike
#pragma unroll
                for (int k = 0; k < C::KP / 2; ++k) {
                  const uint64_t ah = ad0 + (uint64_t)((dx * 16 + 2 * k * C::A_LBO) >> 4);
                  const uint64_t al = ah + (uint64_t)((C::KP * C::A_LBO) >> 4);
                  const uint64_t bh = bd0 + (uint64_t)(((dy * 3 + dx) * C::W_TAP_BYTES + 2 * k * C::B_LBO) >> 4);
                  const uint64_t bl = bh + (uint64_t)((C::KP * C::B_LBO) >> 4);
                  if (elect_one()) {
                    umma_2sm(d, ah, bh, IDESC, (dy | dx | k) != 0);
                    umma_2sm(d + C::CORR_OFF, al, bh, IDESC, (dy | dx | k) != 0);
                    umma_2sm(d + C::CORR_OFF, ah, bl, IDESC, 1);
                  }
                }
              }
            }
          }
          __syncwarp();
          if (elect_one()) {
            umma_commit_both(tfull + a);                                         // accumulator ready -> both epilogues
            umma_commit_both(empty + (it_base + j) % NSLOT);                     // input row j is dead
            if (j == ROW_BLOCK - 1) {
              umma_commit_both(empty + (it_base + ROW_BLOCK) % NSLOT);
              umma_commit_both(empty + (it_base + ROW_BLOCK + 1) % NSLOT);
            }
          }
          __syncwarp();
        }
        it_base += ROW_BLOCK + 2;
      }
    }
  } else if (warp - 2 < EPI_WARPS_USED) {
    // ===== epilogue: TMEM -> registers -> bias / ReLU / mask -> bf16 -> global (channel-group-major, padded rows) ==========
    const int quad = warp & 3;                                 // TMEM lane quadrant this warp may read
    const int part = (warp - 2) >> 2;                          // which half of the output channels
    const int m = quad * 32 + lane;                            // pixel of the tile
    const int cbase = part * C::EPI_COLS;
    const int Wp = row_pitch(P.W);
    uint32_t acc_it = 0;
    for (int u = cluster_id; u < n_units; u += n_clusters) {
      const Unit t = unit_of(P, u, rank);
      const int x = t.x0 + m;
      for (int j = 0; j < ROW_BLOCK; ++j, ++acc_it) {
        const uint32_t a = acc_it % C::NACC, aph = (acc_it / C::NACC) & 1;
        const int y = t.y0 + j;
        [[maybe_unused]] float pv[SPLIT ? C::EPI_COLS : 1];
        if constexpr (SPLIT) {
          // the previous K-half's partial sums of this row are requested BEFORE the wait for the accumulator: their DRAM latency
          // hides behind the row's MMAs
          if (P.pin && t.n < P.N && x < P.W && y < P.H) {
#pragma unroll
            for (int g = 0; g < C::EPI_COLS / 8; ++g) {
              const size_t o32 = ((((size_t)t.n * (COUT / 8) + cbase / 8 + g) * P.H + y) * Wp + x + 1) * 8;
              *reinterpret_cast<float4*>(&pv[g * 8]) = *reinterpret_cast<const float4*>(P.pin + o32);
              *reinterpret_cast<float4*>(&pv[g * 8 + 4]) = *reinterpret_cast<const float4*>(P.pin + o32 + 4);
            }
          }
        }
        mbar_wait(tfull + a, aph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + a * C::ACC_STRIDE + cbase + ((uint32_t)(quad * 32) << 16);
        uint32_t v[C::EPI_COLS];
#pragma unroll
        for (int c0 = 0; c0 < C::EPI_COLS; c0 += 16) tmem_ld16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&v[c0]));
        if constexpr (!SPLIT) {
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(tempty + a);         // accumulator `a` is in registers: the MMA thread may reuse it
          if (t.n < P.N && x < P.W && y < P.H) {
#pragma unroll
            for (int g = 0; g < C::EPI_COLS / 8; ++g) {          // channel groups of 8
              const int cg = cbase / 8 + g;
              const size_t o = ((((size_t)t.n * (COUT / 8) + cg) * P.H + y) * Wp + x + 1) * 8;
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                f[e] = __uint_as_float(v[g * 8 + e]) + P.bias[cbase + g * 8 + e];
                if (P.relu) f[e] = fmaxf(f[e], 0.f);
              }
              if (P.mask) {
                const uint4 mk = *reinterpret_cast<const uint4*>(P.mask + o);
                const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&mk);
#pragma unroll
                for (int e = 0; e < 8; ++e) if (!(__bfloat162float(mb[e]) > 0.f)) f[e] = 0.f;
              }
              uint4 pk;
              __nv_bfloat162* pb = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
              for (int e = 0; e < 4; ++e) pb[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
              *reinterpret_cast<uint4*>(P.out + o) = pk;
            }
          }
        } else {
          uint32_t vc[C::EPI_COLS];                              // correction accumulator (carries a factor 2^11)
#pragma unroll
          for (int c0 = 0; c0 < C::EPI_COLS; c0 += 16) tmem_ld16(taddr + C::CORR_OFF + c0, *reinterpret_cast<uint32_t(*)[16]>(&vc[c0]));
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(tempty + a);
          if (t.n < P.N && x < P.W && y < P.H) {
#pragma unroll
            for (int g = 0; g < C::EPI_COLS / 8; ++g) {
              const int cg = cbase / 8 + g;
              const size_t o32 = ((((size_t)t.n * (COUT / 8) + cg) * P.H + y) * Wp + x + 1) * 8;
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaf(__uint_as_float(vc[g * 8 + e]), 1.f / 2048.f, __uint_as_float(v[g * 8 + e]));
              if (P.pin) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] += pv[g * 8 + e];
              }
              if (P.pout) {                                      // first K-half: hand the partial sum to the second launch
                *reinterpret_cast<float4*>(P.pout + o32) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(P.pout + o32 + 4) = make_float4(f[4], f[5], f[6], f[7]);
                continue;
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                f[e] += P.bias[cbase + g * 8 + e];
                if (P.relu) f[e] = fmaxf(f[e], 0.f);
              }
              // piece planes of channel group cg in a tensor of COUT / 8 groups split into K-halves of out_kp groups
              constexpr int OKP = C::OUT_KP;                       // compile-time: cg is a constant after unrolling
              const int kh = cg / OKP, wi = cg - kh * OKP;
              const size_t plane_hi = (size_t)t.n * (COUT / 4) + (size_t)kh * 2 * OKP + wi;
              const size_t oh = ((plane_hi * P.H + y) * Wp + x + 1) * 8;
              const size_t ol = oh + (size_t)OKP * P.H * Wp * 8;
              if (P.mask16) {
                const uint4 mk = *reinterpret_cast<const uint4*>(P.mask16 + oh);
                const __half* mb = reinterpret_cast<const __half*>(&mk);
#pragma unroll
                for (int e = 0; e < 8; ++e) if (!(__half2float(mb[e]) > 0.f)) f[e] = 0.f;
              }
              if (P.out32) {
                *reinterpret_cast<float4*>(P.out32 + o32) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(P.out32 + o32 + 4) = make_float4(f[4], f[5], f[6], f[7]);
              } else {
                uint4 ph, pl;
                __half2* hh = reinterpret_cast<__half2*>(&ph);
                __half2* hl = reinterpret_cast<__half2*>(&pl);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const __half2 h = cvt_sat_half2(f[2 * e], f[2 * e + 1]);             // saturating: no inf pieces
                  const float2 hf = __half22float2(h);
                  hh[e] = h;
                  hl[e] = cvt_sat_half2((f[2 * e] - hf.x) * 2048.f, (f[2 * e + 1] - hf.y) * 2048.f);
                }
                *reinterpret_cast<uint4*>(P.out16 + oh) = ph;
                *reinterpret_cast<uint4*>(P.out16 + ol) = pl;
              }
            }
          }
        }
      }
    }
  }

  // ---- teardown: nobody may free TMEM (or exit: the peer's MMAs read our shared memory) before both CTAs are done ----------
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_free(tmem_base, C::TMEM_COLS);
}

}  // namespace convtc
}  // namespace dpx
