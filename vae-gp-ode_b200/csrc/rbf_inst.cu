// One translation unit per padded input dimension GPODE_DP (2,4,...,16): instantiates the RBF sweep
// kernels for R = 1 and 2 states per thread and wraps their launches.
#include "rbf_kernels.cuh"

#ifndef GPODE_DP
#error "compile with -DGPODE_DP=<even 2..16>"
#endif

namespace gpode {

namespace {
constexpr int DP = GPODE_DP;

template <typename Kern>
cudaError_t set_smem(Kern kern, int bytes) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

template <typename Kern, typename Args>
cudaError_t launch_sweep(Kern kern, const Args& a, int threads, int R, cudaStream_t st) {
  const int smem = rbf_smem_bytes(a.g);
  cudaError_t e = set_smem(kern, smem);
  if (e != cudaSuccess) return e;
  const long per = static_cast<long>(threads) * R;
  dim3 grid(static_cast<unsigned>((a.g.N + per - 1) / per), static_cast<unsigned>(a.g.L));
  kern<<<grid, threads, smem, st>>>(a);
  return cudaGetLastError();
}
}  // namespace

template <>
cudaError_t rbf_field_fwd_dp<DP>(const RbfFieldFwdArgs& a, cudaStream_t st) {
  int threads, R;
  rbf_pick_shape(a.g, threads, R);
  if (R == 2) return launch_sweep(k_rbf_field_fwd<DP, (DP <= 8 ? 2 : 1)>, a, threads, 2, st);
  return launch_sweep(k_rbf_field_fwd<DP, 1>, a, threads, 1, st);
}

template <>
cudaError_t rbf_field_bwd_dp<DP>(const RbfFieldBwdArgs& a, cudaStream_t st) {
  int threads, R;
  rbf_pick_shape(a.g, threads, R);
  if (R == 2) return launch_sweep(k_rbf_field_bwd<DP, (DP <= 8 ? 2 : 1)>, a, threads, 2, st);
  return launch_sweep(k_rbf_field_bwd<DP, 1>, a, threads, 1, st);
}

template <>
cudaError_t rbf_rollout_fwd_dp<DP>(const RbfRolloutFwdArgs& a, cudaStream_t st) {
  int threads, R;
  rbf_pick_shape(a.g, threads, R);
  if (R == 2) return launch_sweep(k_rbf_rollout_fwd<DP, (DP <= 8 ? 2 : 1)>, a, threads, 2, st);
  return launch_sweep(k_rbf_rollout_fwd<DP, 1>, a, threads, 1, st);
}

template <>
cudaError_t rbf_rollout_bwd_dp<DP>(const RbfRolloutBwdArgs& a, cudaStream_t st) {
  int threads, R;
  rbf_pick_shape(a.g, threads, R);
  if (R == 2) return launch_sweep(k_rbf_rollout_bwd<DP, (DP <= 8 ? 2 : 1)>, a, threads, 2, st);
  return launch_sweep(k_rbf_rollout_bwd<DP, 1>, a, threads, 1, st);
}

template <>
cudaError_t rbf_pgrad_dp<DP>(const RbfPgradArgs& a, cudaStream_t st) {
  const int threads = ((a.g.MP2 + 31) / 32) * 32;
  dim3 grid(static_cast<unsigned>(a.chunks), static_cast<unsigned>(a.g.D_out), static_cast<unsigned>(a.g.L));
  k_rbf_pgrad<DP><<<grid, threads, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gpode
