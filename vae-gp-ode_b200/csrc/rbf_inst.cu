// One translation unit per padded input dimension GPODE_DP (2,4,...,16): instantiates the RBF sweep
// kernels for the register-blocking factors R (states per thread) and wraps their launches.
#include <cstdlib>

#include "rbf_kernels.cuh"
#include "rbf_pgrad_mma.cuh"
#include "rbf_fwd_tc.cuh"
#include "rbf_bwd_tc.cuh"
#include "rbf_small.cuh"

#ifndef GPODE_DP
#error "compile with -DGPODE_DP=<even 2..16>"
#endif

namespace gpode {

namespace {
constexpr int DP = GPODE_DP;
constexpr int RMAX = DP <= 8 ? 4 : 2;

template <typename Kern, typename Args>
cudaError_t launch_sweep(Kern kern, const Args& a, int threads, int R, bool bwd, cudaStream_t st) {
  const int smem = rbf_smem_bytes(a.g, threads, R, bwd);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  const long per = static_cast<long>(threads) * R;
  dim3 grid(static_cast<unsigned>((a.g.N + per - 1) / per), static_cast<unsigned>(a.g.L));
  kern<<<grid, threads, smem, st>>>(a);
  return cudaGetLastError();
}

// picks the instantiation for the heuristic's R (the reverse sweep with R = 4 only exists for DP <= 8)
#define GPODE_DISPATCH_R(KERNEL, a, bwd, st)                                                         \
  int threads, R;                                                                                    \
  rbf_pick_shape((a).g, bwd, threads, R);                                                            \
  if (R >= 4) return launch_sweep(KERNEL<RbfPolicy<DP, ((bwd) ? RMAX : 4)>>, a, threads, ((bwd) ? RMAX : 4), bwd, st); \
  if (R == 2) return launch_sweep(KERNEL<RbfPolicy<DP, 2>>, a, threads, 2, bwd, st);                 \
  return launch_sweep(KERNEL<RbfPolicy<DP, 1>>, a, threads, 1, bwd, st);
}  // namespace

// sweep kernels at D > 8 run on the tensor path (RbfMmaFwdPolicy / RbfMmaBwdPolicy: R = 2 states per thread) once the batch
// fills the chip: measured at config-5 shapes forward 24.8 vs 34.2 ms, reverse sweep 48.2 vs 70.5 ms; a 512-state evaluation
// (the prior at Z of the setup) is latency bound and stays on the FFMA path (0.34 vs 0.58 ms)
template <typename Args, typename KernMma>
cudaError_t launch_fwd_mma(KernMma kern, const Args& a, cudaStream_t st) {
  int threads, R;
  rbf_pick_shape(a.g, false, threads, R, 2);
  return launch_sweep(kern, a, threads, 2, false, st);
}

// tcgen05 forward sweep (rbf_fwd_tc.cuh): 256 states per CTA, one CTA per SM (all 512 tensor-memory columns)
template <typename Args, typename KernTc>
cudaError_t launch_fwd_tc(KernTc kern, const Args& a, cudaStream_t st) {
  const int smem = rbf_fwd_tc_smem_bytes(a.g);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  dim3 grid(static_cast<unsigned>((a.g.N + kFtStates - 1) / kFtStates), static_cast<unsigned>(a.g.L));
  kern<<<grid, kFtThreads, smem, st>>>(a);
  return cudaGetLastError();
}

// small batches (rbf_small.cuh): 32 states per CTA, 16 warps splitting the rows of every output dimension, parameters resident in shared memory
template <typename Args, typename KernSmall>
cudaError_t launch_small(KernSmall kern, const Args& a, cudaStream_t st) {
  const int smem = rbf_small_smem_bytes(a.g);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  // the outputs of a state block are split over a thread-block cluster along z when the launch leaves SMs idle (rbf_small.cuh)
  const int C = rbf_small_cluster(a.g);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>((a.g.N + kSmStates - 1) / kSmStates), static_cast<unsigned>(a.g.L), static_cast<unsigned>(C));
  cfg.blockDim = dim3(kSmThreads);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = static_cast<unsigned>(C);
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e2 = cudaLaunchKernelEx(&cfg, kern, a);
  if (e2 != cudaSuccess && C > 1) {   // a cluster of this size cannot be placed (partitioned device, ...): the same kernel without the split
    (void)cudaGetLastError();
    cfg.gridDim.z = 1;
    attr[0].val.clusterDim.z = 1;
    e2 = cudaLaunchKernelEx(&cfg, kern, a);
  }
  return e2;
}

template <>
cudaError_t rbf_field_fwd_dp<DP>(const RbfFieldFwdArgs& a, cudaStream_t st) {
  if (rbf_use_small(a.g)) return launch_small(k_field_fwd<RbfSmallPolicy<DP>>, a, st);
  if constexpr (DP > 8) {   // (measured at D = 6: 1.17 vs 1.22 ms forward, 2.12 vs 2.10 ms reverse -- no gain below D = 9)
    if (rbf_fwd_use_tc(a.g)) return launch_fwd_tc(k_field_fwd<RbfTcFwdPolicy<DP>>, a, st);
    if (rbf_fwd_use_mma(a.g)) return launch_fwd_mma(k_field_fwd<RbfMmaFwdPolicy<DP>>, a, st);
  }
  GPODE_DISPATCH_R(k_field_fwd, a, false, st)
}
template <typename Args, typename KernMma>
cudaError_t launch_bwd_mma(KernMma kern, const Args& a, cudaStream_t st) {
  int threads, R;
  rbf_pick_shape(a.g, true, threads, R, 2);
  return launch_sweep(kern, a, threads, 2, true, st);
}

// tcgen05 reverse sweep (rbf_bwd_tc.cuh): 128 states per CTA, one CTA per SM (all 512 tensor-memory columns)
template <typename Args, typename KernTc>
cudaError_t launch_bwd_tc(KernTc kern, const Args& a, cudaStream_t st) {
  const int smem = rbf_bwd_tc_smem_bytes(a.g);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  dim3 grid(static_cast<unsigned>((a.g.N + kBtStates - 1) / kBtStates), static_cast<unsigned>(a.g.L));
  kern<<<grid, kBtThreads, smem, st>>>(a);
  return cudaGetLastError();
}

template <>
cudaError_t rbf_field_bwd_dp<DP>(const RbfFieldBwdArgs& a, cudaStream_t st) {
  if (rbf_use_small(a.g)) return launch_small(k_field_bwd<RbfSmallPolicy<DP>>, a, st);
  if constexpr (DP > 8) {   // (measured at D = 6: 1.17 vs 1.22 ms forward, 2.12 vs 2.10 ms reverse -- no gain below D = 9)
    if (rbf_bwd_use_tc(a.g)) return launch_bwd_tc(k_field_bwd<RbfTcBwdPolicy<DP>>, a, st);
    if (rbf_fwd_use_mma(a.g)) return launch_bwd_mma(k_field_bwd<RbfMmaBwdPolicy<DP>>, a, st);
  }
  GPODE_DISPATCH_R(k_field_bwd, a, true, st)
}
template <>
cudaError_t rbf_rollout_fwd_dp<DP>(const RbfRolloutFwdArgs& a, cudaStream_t st) {
  if (rbf_use_small(a.g)) return launch_small(k_rollout_fwd<RbfSmallPolicy<DP>>, a, st);
  if constexpr (DP > 8) {   // (measured at D = 6: 1.17 vs 1.22 ms forward, 2.12 vs 2.10 ms reverse -- no gain below D = 9)
    if (rbf_fwd_use_tc(a.g)) return launch_fwd_tc(k_rollout_fwd<RbfTcFwdPolicy<DP>>, a, st);
    if (rbf_fwd_use_mma(a.g)) return launch_fwd_mma(k_rollout_fwd<RbfMmaFwdPolicy<DP>>, a, st);
  }
  GPODE_DISPATCH_R(k_rollout_fwd, a, false, st)
}
template <>
cudaError_t rbf_rollout_bwd_dp<DP>(const RbfRolloutBwdArgs& a, cudaStream_t st) {
  if (rbf_use_small(a.g)) return launch_small(k_rollout_bwd<RbfSmallPolicy<DP>>, a, st);
  if constexpr (DP > 8) {   // (measured at D = 6: 1.17 vs 1.22 ms forward, 2.12 vs 2.10 ms reverse -- no gain below D = 9)
    if (rbf_bwd_use_tc(a.g)) return launch_bwd_tc(k_rollout_bwd<RbfTcBwdPolicy<DP>>, a, st);
    if (rbf_fwd_use_mma(a.g)) return launch_bwd_mma(k_rollout_bwd<RbfMmaBwdPolicy<DP>>, a, st);
  }
  GPODE_DISPATCH_R(k_rollout_bwd, a, true, st)
}

namespace {
// D > 8: tensor-path kernels (3xTF32);  D <= 8: FFMA kernel
template <int D>
cudaError_t launch_pgrad(const RbfPgradArgs& a, cudaStream_t st) {
  if constexpr (D > 8) {
    // warp-level tensor path (mma.sync): each warp owns 16 * MT inducing points, a CTA 128 * MT
    int MT, n_mblk;
    rbf_pgrad_mma_shape(a.g, MT, n_mblk);
    dim3 grid(static_cast<unsigned>(a.chunks * n_mblk), static_cast<unsigned>(a.g.D_out), static_cast<unsigned>(a.g.L));
    const int smem = rbf_pgrad_mma_smem_bytes(2);
    cudaError_t e = cudaFuncSetAttribute(MT == 2 ? k_rbf_pgrad_mma<2, 2> : k_rbf_pgrad_mma<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    if (MT == 2) k_rbf_pgrad_mma<2, 2><<<grid, kPgmThreads, smem, st>>>(a);
    else k_rbf_pgrad_mma<2, 1><<<grid, kPgmThreads, smem, st>>>(a);
  } else {
    int threads, PP, n_mblk;
    rbf_pgrad_shape(a.g, threads, PP, n_mblk);
    dim3 grid(static_cast<unsigned>(a.chunks * n_mblk), static_cast<unsigned>(a.g.D_out), static_cast<unsigned>(a.g.L));
    k_rbf_pgrad<D, 1><<<grid, threads, 0, st>>>(a);
  }
  return cudaGetLastError();
}
}  // namespace

template <>
cudaError_t rbf_pgrad_dp<DP>(const RbfPgradArgs& a, cudaStream_t st) { return launch_pgrad<DP>(a, st); }

}  // namespace gpode
