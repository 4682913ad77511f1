// dpx_conv_umma.cuh — 3x3 / stride 1 / pad 1 convolution as an implicit GEMM on the 5th-generation tensor cores:
// tcgen05.mma (UTCHMMA) with TMEM accumulators, operands staged by TMA (im2col mode for the NHWC activations,
// tiled mode for the KRSC filters), bias + optional ReLU fused in the TMEM->register epilogue, TMA store.
// Instantiated from the CUTLASS/CuTe sm100 collectives (header-only, vendored with flashinfer) — the FFDNet
// denoiser behind `deep_prior` (proxfn/pnp/denoisers/models/network_ffdnet.py:44-52) is the only GEMM-shaped op on
// the proximal-iteration path.
#pragma once
#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/kernel_hardware_info.hpp"
#include "cutlass/conv/convolution.h"
#include "cutlass/conv/convnd_problem_shape.hpp"
#include "cutlass/conv/dispatch_policy.hpp"
#include "cutlass/conv/collective/collective_builder.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/conv/device/conv_universal_adapter.hpp"
#include "cutlass/conv/kernel/conv_universal.hpp"
#include "cutlass/epilogue/fusion/operations.hpp"
#include "cutlass/epilogue/thread/activation.h"
#include "cutlass/util/packed_stride.hpp"

#include "dpx_conv_api.h"

namespace dpx {
namespace conv {

using namespace cute;

// TILE_N = output channels per MMA tile, TILE_K = input channels per k-block, RELU = fuse max(.,0),
// TWO_SM = pair two SMs on one 256-row tile (tcgen05 cta_group::2): the filter tile is fetched once per pair
template <int TILE_N, int TILE_K, bool RELU, bool TWO_SM = false>
struct Conv3x3 {
  using ElementAct = cutlass::bfloat16_t;
  using ElementFlt = cutlass::bfloat16_t;
  using ElementOut = cutlass::bfloat16_t;
  using ElementAcc = float;
  using ElementCompute = float;
  using ElementBias = float;
  static constexpr int Align = 8;                                  // 16-byte TMA alignment of the channel axis
  static constexpr cutlass::conv::Operator ConvOp = cutlass::conv::Operator::kFprop;
  using MmaTileShape = Shape<Int<TWO_SM ? 256 : 128>, Int<TILE_N>, Shape<Int<TILE_K>>>;
  using ClusterShape = cute::conditional_t<TWO_SM, Shape<_2, _1, _1>, Shape<_1, _1, _1>>;
  using KernelSchedule = cute::conditional_t<TWO_SM, cutlass::conv::KernelImplicitTmaWarpSpecialized2SmSm100,
                                             cutlass::conv::KernelImplicitTmaWarpSpecialized1SmSm100>;
  using EpilogueSchedule = cute::conditional_t<TWO_SM, cutlass::epilogue::TmaWarpSpecialized2Sm, cutlass::epilogue::TmaWarpSpecialized1Sm>;
  template <class T> using Act = cute::conditional_t<RELU, cutlass::epilogue::thread::ReLu<T>, cutlass::epilogue::thread::Identity<T>>;
  using FusionOp = cutlass::epilogue::fusion::LinCombPerColBiasEltAct<Act, ElementOut, ElementCompute, ElementBias>;

  using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, MmaTileShape, ClusterShape,
      cutlass::epilogue::collective::EpilogueTileAuto, ElementAcc, ElementCompute,
      ElementOut, cutlass::layout::TensorNHWC, Align, ElementOut, cutlass::layout::TensorNHWC, Align,
      EpilogueSchedule, FusionOp>::CollectiveOp;
  using CollectiveMainloop = typename cutlass::conv::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, ConvOp,
      ElementAct, cutlass::layout::TensorNHWC, Align, ElementFlt, cutlass::layout::TensorNHWC, Align,
      ElementAcc, MmaTileShape, ClusterShape,
      cutlass::conv::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
      KernelSchedule>::CollectiveOp;
  using ProblemShape = cutlass::conv::ConvProblemShape<ConvOp, CollectiveMainloop::DispatchPolicy::NumSpatialDimensions>;
  using ConvKernel = cutlass::conv::kernel::ConvUniversal<ProblemShape, CollectiveMainloop, CollectiveEpilogue>;
  using Conv = cutlass::conv::device::ConvUniversalAdapter<ConvKernel>;

  // act [n,h,w,c] bf16, flt [k,3,3,c] bf16, bias [k] fp32, out [n,h,w,k] bf16; returns 0 on success
  static int run(const void* act, const void* flt, const float* bias, void* out, int n, int h, int w, int c, int k,
                 void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    ProblemShape ps(cutlass::conv::Mode::kCrossCorrelation, {n, h, w, c}, {k, 3, 3, c}, {1, 1}, {1, 1}, {1, 1}, {1, 1}, 1);
    using StrideC = typename ConvKernel::StrideC;
    using StrideD = typename ConvKernel::StrideD;
    // output strides ((q, p, n), _1) of the packed NHWC tensor
    StrideC sc{};
    StrideD sd{};
    cute::for_each(cute::make_seq<cute::rank<0>(StrideC{})>{}, [&](auto i) {
      cute::get<0, i>(sc) = ps.stride_C[ProblemShape::RankT - 2 - i];
      cute::get<0, i>(sd) = ps.stride_C[ProblemShape::RankT - 2 - i];
    });
    typename Conv::Arguments args{ps, {(const ElementAct*)act, (const ElementFlt*)flt}, {{}, nullptr, sc, (ElementOut*)out, sd}};
    args.epilogue.thread.alpha = 1.f;
    args.epilogue.thread.beta = 0.f;
    args.epilogue.thread.bias_ptr = bias;
    Conv op;
    if (op.can_implement(args) != cutlass::Status::kSuccess) return 1;
    if (Conv::get_workspace_size(args) > workspace_bytes) return 4;
    if (op.initialize(args, workspace, stream) != cutlass::Status::kSuccess) return 2;
    if (op.run(stream) != cutlass::Status::kSuccess) return 3;
    return 0;
  }
  static size_t workspace_size(int n, int h, int w, int c, int k) {
    ProblemShape ps(cutlass::conv::Mode::kCrossCorrelation, {n, h, w, c}, {k, 3, 3, c}, {1, 1}, {1, 1}, {1, 1}, {1, 1}, 1);
    typename Conv::Arguments args{ps, {nullptr, nullptr}, {{}, nullptr, {}, nullptr, {}}};
    return Conv::get_workspace_size(args);
  }
};

}  // namespace conv
}  // namespace dpx
