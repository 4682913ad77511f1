// dpx_conv_wgrad.cuh — weight gradient of the 3x3 / stride 1 / pad 1 convolution on the 5th-generation tensor cores (sm_100a):
//     dW[tap][co][ci] = sum over (image, y, x) of  gy[co][y][x] * a[ci][y + ky - 1][x + kx - 1],      tap = ky * 3 + kx
// (torch: conv2d's grad_weight; the layers of network_ffdnet.py:27-68 under training).
//
// Formulation.  A GEMM whose REDUCTION dimension is the pixel: per 128-pixel piece of an image row and per tap
//     D_tap[co, ci] += GY[pixel, co]^T  A_tap[pixel, ci]          (M = 128 >= C_out, N = C_in, K = 128 pixels = 8 MMAs of K = 16).
// Both operands are the channel-group-major activation tensors the forward kernel reads and writes, [N][C/8][h][row_pitch(W)][8] bf16:
// for one pixel the 8 channels of a group are 16 contiguous bytes and consecutive pixels follow at 16 bytes -- exactly the
// canonical MN-MAJOR shared-memory operand of tcgen05.mma without swizzle (an 8 pixel x 8 channel core matrix = 128 contiguous
// bytes; next 8 pixels +128 B = the descriptor's leading-dimension byte offset; next channel group + one staged row of a group =
// its stride byte offset; cute/atom/mma_traits_sm100.hpp, "make_umma_desc<Major::MN>").  So the rows are staged by the SAME TMA
// boxes as in the forward kernel (130 pixels x all channel groups per row), a tap is again just a start address (+16 B per
// pixel of horizontal shift, another ring slot per vertical shift), and no transposition exists anywhere.
//   * accumulators: one 128 x C_in fp32 tile per tap in TMEM, alive for the whole kernel (never read back until the end).  Nine
//     taps x 96 columns exceed the 512 TMEM columns, so blockIdx.y splits the taps (0..4 / 5..8) over two CTAs that stream the
//     same rows (L2 hits).  M = 128 reads 16 channel groups of GY where only C_out / 8 exist: the rows of D beyond C_out are
//     garbage and ignored (each row of D depends on its own row of GY^T only); the gy ring is followed by the a ring, so the
//     over-read stays inside initialised shared memory.
//   * at the end each CTA adds its tiles to dW (fp32, global atomics: 148 x 2 CTAs x 9 x 96 x 96 values).
// Every staged pixel enters the sum, so a run must never reach into the next row: the padded rows are extended with zeros to a
// whole number of 128-pixel tiles (row_pitch), and pixels beyond the image width are never written by any producer.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocation + MMA issue, warps 2..5 = final read-out.
#pragma once
#include "dpx_conv_tc.cuh"

namespace dpx {
namespace convtc {

constexpr int WG_NA = 5;                  // activation row slots (3 live + 2 in flight)
constexpr int WG_NG = 3;                  // output-gradient row slots
constexpr int WG_THREADS = 192;

template <int CGG, int CGA>               // channel groups of gy (C_out / 8) and of a (C_in / 8)
struct WgCfg {
  static constexpr int N = CGA * 8;                                       // UMMA N = input channels
  static constexpr uint32_t G_ROW_BYTES = CGG * HALO_PX * 16, A_ROW_BYTES = CGA * HALO_PX * 16;
  static constexpr uint32_t G_SLOT = (G_ROW_BYTES + 127) / 128 * 128, A_SLOT = (A_ROW_BYTES + 127) / 128 * 128;
  static constexpr uint32_t GROUP_STRIDE = HALO_PX * 16;                  // one channel group of a staged row
  // M = 128 reads 16 groups of gy starting inside the gy ring: the ring + the a ring behind it must cover the over-read
  static constexpr size_t SMEM = 1024 + (size_t)WG_NG * G_SLOT + (size_t)WG_NA * A_SLOT + 256;
  static_assert((size_t)WG_NA * A_SLOT >= (size_t)(16 - CGG) * GROUP_STRIDE, "the a ring must cover the over-read of the last gy slot");
  static constexpr int COLS_PER_TAP = N < 32 ? 32 : N;                    // TMEM columns per tap accumulator
  static constexpr int MAX_TAPS = 5;
  static constexpr int TMEM_COLS = 512;
  static_assert(COLS_PER_TAP * MAX_TAPS <= 512 && N % 16 == 0 && N <= 256, "shape");
};

struct WgParams {
  float* dw;                       // [9][128][N] fp32, accumulated with atomics (rows >= C_out are never written)
  int cout;                        // valid rows of D
  int N, H, W;
  int n_tiles, x_tiles, row_blocks;
};

// one staged row: same box as the forward kernel, completion on this CTA's barrier (single-CTA form)
__device__ __forceinline__ void tma_row1(void* dst, const CUtensorMap* map, uint64_t* bar, int xs, int y, int plane0) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(s2u(dst)), "l"(map), "r"(s2u(bar)), "r"(0), "r"(0), "r"(xs), "r"(y), "r"(plane0)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc1(uint32_t* dst_smem, int cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2u(dst_smem)), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish1() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_free1(uint32_t addr, int cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], M = 128 on one CTA, K = 16 bf16
__device__ __forceinline__ void umma_1sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit1(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2u(b)) : "memory");
}
// instruction descriptor of kind::f16 with BOTH operands MN-major (bits 15 / 16), bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t instr_desc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int CGG, int CGA>
__global__ void __launch_bounds__(WG_THREADS, 1)
    k_conv3x3_wgrad(const __grid_constant__ CUtensorMap gy_map, const __grid_constant__ CUtensorMap a_map, WgParams P) {
  using C = WgCfg<CGG, CGA>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sm_g = base;                                               // gy ring
  uint8_t* sm_a = base + (size_t)WG_NG * C::G_SLOT;                   // a ring (also absorbs the M = 128 over-read of the gy ring)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_a + (size_t)WG_NA * C::A_SLOT);
  uint64_t* full_a = bars;                                            // [WG_NA]
  uint64_t* empty_a = bars + WG_NA;
  uint64_t* full_g = bars + 2 * WG_NA;                                // [WG_NG]
  uint64_t* empty_g = bars + 2 * WG_NA + WG_NG;
  uint64_t* done = bars + 2 * WG_NA + 2 * WG_NG;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap0 = blockIdx.y == 0 ? 0 : C::MAX_TAPS, ntap = blockIdx.y == 0 ? C::MAX_TAPS : 9 - C::MAX_TAPS;
  const int cta = blockIdx.x, n_cta = gridDim.x;

  // the over-read of M = 128 past the last staged group must see finite numbers: clear everything once
  for (uint32_t o = threadIdx.x * 16; o < (uint32_t)(WG_NG * C::G_SLOT + WG_NA * C::A_SLOT); o += WG_THREADS * 16)
    *reinterpret_cast<uint4*>(base + o) = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_NA; ++i) { mbar_init(full_a + i, 1); mbar_init(empty_a + i, 1); }
    for (int i = 0; i < WG_NG; ++i) { mbar_init(full_g + i, 1); mbar_init(empty_g + i, 1); }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy zeros before the async-proxy (TMA, MMA) uses
  __syncthreads();
  if (warp == 1) {
    tmem_alloc1(tmem_slot, C::TMEM_COLS);
    tmem_relinquish1();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&gy_map) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a_map) : "memory");
      uint32_t ia = 0, ig = 0;
      for (int u = cta; u < P.n_tiles; u += n_cta) {
        const int xt = u % P.x_tiles, rb = (u / P.x_tiles) % P.row_blocks, n = u / (P.x_tiles * P.row_blocks);
        const int x0 = xt * TILE_PX, y0 = rb * ROW_BLOCK;
        for (int r = 0; r < ROW_BLOCK + 2; ++r, ++ia) {
          const uint32_t s = ia % WG_NA, ph = (ia / WG_NA) & 1;
          mbar_wait(empty_a + s, ph ^ 1);
          mbar_expect_tx(full_a + s, C::A_ROW_BYTES);
          tma_row1(sm_a + (size_t)s * C::A_SLOT, &a_map, full_a + s, x0, y0 - 1 + r, n * CGA);
          if (r >= 2) {                                               // the gy row of output row r - 2 goes out with its last a row
            const uint32_t sg = ig % WG_NG, pg = (ig / WG_NG) & 1;
            mbar_wait(empty_g + sg, pg ^ 1);
            mbar_expect_tx(full_g + sg, C::G_ROW_BYTES);
            tma_row1(sm_g + (size_t)sg * C::G_SLOT, &gy_map, full_g + sg, x0, y0 + r - 2, n * CGG);
            ++ig;
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t IDESC = instr_desc_mn(128, C::N);
    const uint32_t a0 = s2u(sm_a), g0 = s2u(sm_g);
    uint32_t ia_base = 0, ig = 0;
    uint32_t first = 1;
    for (int u = cta; u < P.n_tiles; u += n_cta) {
      int waited = 0;
      for (int j = 0; j < ROW_BLOCK; ++j, ++ig) {
        while (waited < j + 3) {
          const uint32_t i = ia_base + waited;
          mbar_wait(full_a + (i % WG_NA), (i / WG_NA) & 1);
          ++waited;
        }
        const uint32_t sg = ig % WG_NG;
        mbar_wait(full_g + sg, (ig / WG_NG) & 1);
        tc_fence_after();
        // MN-major, no swizzle: leading byte offset = next 8 pixels (128 B), stride byte offset = next channel group.
        // pixel x0 + i of the tile sits at padded index i + 1 of the staged gy row; tap (dy, dx) reads a at padded index i + dx
        const uint64_t gd0 = smem_desc(g0 + sg * C::G_SLOT + 16, 128, C::GROUP_STRIDE);
        for (int tp = 0; tp < ntap; ++tp) {
          const int tap = tap0 + tp, dy = tap / 3, dx = tap - 3 * dy;
          const uint32_t slot = (ia_base + j + dy) % WG_NA;
          const uint64_t ad0 = smem_desc(a0 + slot * C::A_SLOT + dx * 16, 128, C::GROUP_STRIDE);
          const uint32_t d = tmem_base + tp * C::COLS_PER_TAP;
#pragma unroll
          for (int kk = 0; kk < TILE_PX / 16; ++kk) {
            const uint64_t gd = gd0 + (uint64_t)((kk * 256) >> 4), ad = ad0 + (uint64_t)((kk * 256) >> 4);
            if (elect_one()) umma_1sm(d, gd, ad, IDESC, (first == 0 || kk != 0) ? 1u : 0u);
          }
        }
        first = 0;
        __syncwarp();
        if (elect_one()) {
          umma_commit1(empty_g + sg);
          umma_commit1(empty_a + (ia_base + j) % WG_NA);
          if (j == ROW_BLOCK - 1) {
            umma_commit1(empty_a + (ia_base + ROW_BLOCK) % WG_NA);
            umma_commit1(empty_a + (ia_base + ROW_BLOCK + 1) % WG_NA);
          }
        }
        __syncwarp();
      }
      ia_base += ROW_BLOCK + 2;
    }
    __syncwarp();
    if (elect_one()) umma_commit1(done);                             // every MMA of this CTA has completed
    __syncwarp();
  } else {
    // ===== read-out: D_tap rows (= output channels) live in the TMEM lanes, columns = input channels =====================
    const bool has_work = cta < P.n_tiles;
    mbar_wait(done, 0);
    tc_fence_after();
    const int quad = warp & 3;
    const int co = quad * 32 + lane;
    if (has_work) {
      for (int tp = 0; tp < ntap; ++tp) {
#pragma unroll 1
        for (int c0 = 0; c0 < C::N; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + tp * C::COLS_PER_TAP + c0 + ((uint32_t)(quad * 32) << 16), v);
          tmem_ld_wait();
          if (co < P.cout) {
            float* dst = P.dw + ((size_t)(tap0 + tp) * 128 + co) * C::N + c0;
#pragma unroll
            for (int e = 0; e < 16; ++e) atomicAdd(dst + e, __uint_as_float(v[e]));
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_free1(tmem_base, C::TMEM_COLS);
}

}  // namespace convtc
}  // namespace dpx
