// Host-side descriptors of the per-rollout setup kernels (setup_kernels.cu): K(Z,Z) (RBF or divergence-free) + Cholesky + whitened solves
// (compute_nu), inducing sample and KL on the packed lower-triangular parameter.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace gpode {

struct NuGeom {
  int L, M, D_in, D_out;
  int dimwise;     // 1: one (M x M) system per output dim (RBF dimwise); 0: one system shared by all output dims (RBF shared)
  int df;          // 1: divergence-free kernel -- ONE (M D x M D) system shared by the L samples (core/kernels.py:289-303,376-387)
  int n;           // order of each matrix: M (RBF) or M * D (DF)
  int Kc;          // number of matrices: D_out (RBF dimwise) or 1
  int NR;          // right-hand-side columns per matrix: L (RBF dimwise, DF) or L * D_out (RBF shared)
  float jitter;    // 1e-5 (core/kernels.py:11)
};

size_t nu_save_floats(const NuGeom& g);   // Cholesky factors (Kc,M,M) + Lc^-1 u_prior (Kc,M,NR), forward -> backward
size_t nu_ws_floats(const NuGeom& g);
cudaError_t nu_forward(const NuGeom& g, const float* Z, const float* ell, const float* var, const float* u_prior, const float* u, float* nu,
                       float* save, int* info, float* ws, cudaStream_t st);
cudaError_t nu_backward(const NuGeom& g, const float* Z, const float* ell, const float* var, const float* u, const float* save, const float* dnu,
                        float* d_uprior, float* d_u, float* d_Z, float* d_ell, float* d_var, float* ws, cudaStream_t st);
cudaError_t inducing_forward(int L, int M, int D, const float* Lq, const float* Um, const float* eps, float* u, cudaStream_t st);
cudaError_t inducing_backward(int L, int M, int D, const float* eps, const float* du, float* dLq, float* dUm, int accumulate, cudaStream_t st);
cudaError_t kl_forward(int M, int D, const float* Lq, const float* Um, float* kl, cudaStream_t st);
cudaError_t kl_backward(int M, int D, const float* Lq, const float* Um, const float* dkl, float* dLq, float* dUm, cudaStream_t st);

}  // namespace gpode
