// One translation unit per DF dimension GPODE_DF_D (1..8): instantiates the sweep kernels with the DF
// policy for the register-blocking factors R and wraps their launches.
#include "df_kernels.cuh"

#ifndef GPODE_DF_D
#error "compile with -DGPODE_DF_D=<1..8>"
#endif

namespace gpode {

namespace {
constexpr int D = GPODE_DF_D;

template <typename Kern, typename Args>
cudaError_t launch_sweep(Kern kern, const Args& a, int threads, int R, bool bwd, cudaStream_t st, int W = 0) {
  const int smem = df_smem_bytes(a.g, threads, R, bwd, true, W);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  const long per = W > 0 ? 32 : static_cast<long>(threads) * R;
  const int C = df_cluster(a.g, static_cast<int>(per));   // small batches: the rows of every chunk are split over a thread-block cluster along z
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>((a.g.N + per - 1) / per), static_cast<unsigned>(a.g.L), static_cast<unsigned>(C));
  cfg.blockDim = dim3(static_cast<unsigned>(threads));
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
    (void)cudaGetLastError();         // (the exchange buffers sized for C stay allocated; gridDim.z = 1 makes the kernels skip them)
    cfg.gridDim.z = 1;
    attr[0].val.clusterDim.z = 1;
    e2 = cudaLaunchKernelEx(&cfg, kern, a);
  }
  return e2;
}

#define GPODE_DF_DISPATCH_R(KERNEL, a, bwd, st)                                            \
  if (df_use_small((a).g)) return launch_sweep(KERNEL<DfPolicy<D, 1, kDfSmallW>>, a, 32 * kDfSmallW, 1, bwd, st, kDfSmallW); \
  int threads, R;                                                                          \
  df_pick_shape((a).g, bwd, threads, R);                                                   \
  if (R == 2) return launch_sweep(KERNEL<DfPolicy<D, 2>>, a, threads, 2, bwd, st);         \
  return launch_sweep(KERNEL<DfPolicy<D, 1>>, a, threads, 1, bwd, st);
}  // namespace

template <>
cudaError_t df_field_fwd_d<D>(const DfFieldFwdArgs& a, cudaStream_t st) { GPODE_DF_DISPATCH_R(k_field_fwd, a, false, st) }
template <>
cudaError_t df_rollout_fwd_d<D>(const DfRolloutFwdArgs& a, cudaStream_t st) { GPODE_DF_DISPATCH_R(k_rollout_fwd, a, false, st) }
template <>
cudaError_t df_field_bwd_d<D>(const DfFieldBwdArgs& a, cudaStream_t st) {
  if (df_use_small(a.g)) return launch_sweep(k_field_bwd<DfPolicy<D, 1, kDfSmallW>>, a, 32 * kDfSmallW, 1, true, st, kDfSmallW);
  return launch_sweep(k_field_bwd<DfPolicy<D, 1>>, a, [&] { int t, r; df_pick_shape(a.g, true, t, r); return t; }(), 1, true, st);
}
template <>
cudaError_t df_rollout_bwd_d<D>(const DfRolloutBwdArgs& a, cudaStream_t st) {
  if (df_use_small(a.g)) return launch_sweep(k_rollout_bwd<DfPolicy<D, 1, kDfSmallW>>, a, 32 * kDfSmallW, 1, true, st, kDfSmallW);
  return launch_sweep(k_rollout_bwd<DfPolicy<D, 1>>, a, [&] { int t, r; df_pick_shape(a.g, true, t, r); return t; }(), 1, true, st);
}

template <>
cudaError_t df_pgrad_d<D>(const DfPgradArgs& a, cudaStream_t st) {
  // the inducing-point gradients (dnu, dZ, dc) come out of the reverse sweep itself; only the feature operator B is left
  dim3 gb(static_cast<unsigned>(a.chunks_b), static_cast<unsigned>((a.g.D * a.g.SP2 + kDfPgThreads - 1) / kDfPgThreads), static_cast<unsigned>(a.g.L));
  k_df_pgrad<D><<<gb, kDfPgThreads, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gpode
