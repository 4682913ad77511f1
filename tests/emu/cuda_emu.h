// cuda_emu.h — TEST INFRASTRUCTURE: a tiny CUDA-thread emulator so that the fused FFT kernels
// (delta-prox_b200/csrc/dpx_fused_kernels.cuh) compile with g++ and run on the CPU box, one OS thread per
// CUDA thread of a block, __syncthreads() mapped to a std::barrier.  Only what those kernels use exists.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <barrier>
#include <functional>
#include <thread>
#include <vector>

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

namespace emu {
inline thread_local dim3 t_threadIdx, t_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline std::barrier<>* g_barrier = nullptr;
inline std::vector<unsigned char> g_smem;

// run `kernel()` for every block of the grid; blocks are executed one after the other
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& kernel) {
  g_blockDim = block; g_gridDim = grid;
  g_smem.assign(smem_bytes + 64, 0);
  std::barrier<> bar((std::ptrdiff_t)block.x);
  g_barrier = &bar;
  std::vector<std::thread> th;
  for (unsigned t = 0; t < block.x; ++t) {
    th.emplace_back([&, t]() {
      for (unsigned bz = 0; bz < grid.z; ++bz)
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          t_threadIdx = dim3(t); t_blockIdx = dim3(bx, by, bz);
          kernel();
          g_barrier->arrive_and_wait();      // block boundary: shared memory is reused by the next block
        }
    });
  }
  for (auto& x : th) x.join();
  g_barrier = nullptr;
}
}  // namespace emu

#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim
#define __syncthreads() emu::g_barrier->arrive_and_wait()
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define DPX_HD inline
#define DPX_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::g_smem.data())
static inline float2 __ldg(const float2* p) { return *p; }
static inline float __ldg(const float* p) { return *p; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
