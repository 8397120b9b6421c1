// Host-side descriptors of the divergence-free (DF) kernels (shared by the kernel translation units and the C ABI).
//
// Math (reference experiments/model/core/kernels.py:201-393), D_in = D_out = D:
//   prior   f_p[c] = sum_{a,s} B'[s,a,c] cos(x . Om'[:,s,a] + b'[s,a])
//             Om'[d,s,a] = eps[d,s,a] / ell[a,d]                                   (kernels.py:120-124)
//             cos(th) w[s,a] + sin(th) w[S+s,a] = R cos(th - phi)  (R = hypot, phi = atan2)  -> one MUFU instead of two
//             b' = b - phi,  B'[s,a,c] = sqrt(var_c / S) R[s,a] B[s,a,c]             (kernels.py:340-349)
//   update  f_u[j] = sum_m [ d_j sum_i nu_mi d_i e_ij + nu_mj e_jj (h_j - r2) ],   d = x - z_m, r2 = |d|^2
//             e_ij = var_j c_ij^2 exp(-r2 c_ij / 2) = 2^(r2 k_ij + lc_ij),  c_ij = 1 / ell_ij^2,
//             k_ij = -log2(e) c_ij / 2,  lc_ij = log2(var_j c_ij^2),  h_j = (D - 1) / c_jj   (kernels.py:289-301,390-393)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "chunk_geom.h"
#include "gpode.h"
#include "sweep_args.h"

namespace gpode {

constexpr int kDfMaxD = 8;

struct DfGeom {
  int L, N, NL;
  int D_in, D_out, D;        // all equal (the reference DF kernel is square)
  int M, S, MP2, SP2;        // pairs (rounded up)
  int NCs, NCm, RCs, RCm;    // chunks / rows per chunk: feature section of one block a, inducing section
  int stage_floats;
  int rowf_s, rowf_m;        // floats per feature row (2D+2 float2) / inducing row (2D float2)
  int hdr_floats;            // header: KC[D*D] float4 {k,k,lc,lc}, then h[D] (padded to 4)
  int order, off;            // 1, 0
  ChunkGeom cg;
};
// packed = [hdr_floats] header (sample independent), then per sample [D][SP2][rowf_s] feature rows + [MP2][rowf_m] inducing rows
inline size_t df_sample_floats(const DfGeom& g) { return static_cast<size_t>(g.D) * g.SP2 * g.rowf_s + static_cast<size_t>(g.MP2) * g.rowf_m; }
inline size_t df_packed_floats(const DfGeom& g) { return g.hdr_floats + static_cast<size_t>(g.L) * df_sample_floats(g); }
__host__ __device__ inline const float* df_rows_ptr(const float* packed, const DfGeom& g, int l) {
  return packed + g.hdr_floats + static_cast<size_t>(l) * (static_cast<size_t>(g.D) * g.SP2 * g.rowf_s + static_cast<size_t>(g.MP2) * g.rowf_m);
}

struct DfAccum {             // fp32 accumulators in the workspace, zeroed before the backward
  float* dnu;                // [L][2*MP2][D]        dL/dnu
  float* dz;                 // [2*MP2][D]           dL/dZ (summed over samples)
  float* dbp;                // [L][D(a)][2*SP2][D(c)]  sum_n g_c cos(theta'_sa)
  float* dc;                 // [D][D]   sum u_j p_i e_ij (r2 k_ij + 2 log2e) + diagonal terms   (dL/dc_ij * c_ij / ln2)
  float* dell_x;             // [D(a)][D(d)]  sum_n x_d q_ad    (theta path of the lengthscales)
  float* dvar;               // [D]      sum_n g (f - f_p / 2)
};
size_t df_acc_floats(const DfGeom& g);
DfAccum df_acc(float* base, const DfGeom& g);

using DfFieldFwdArgs = FieldFwdArgsT<DfGeom>;
using DfRolloutFwdArgs = RolloutFwdArgsT<DfGeom>;
using DfRolloutBwdArgs = RolloutBwdArgsT<DfGeom, DfAccum>;
using DfFieldBwdArgs = FieldBwdArgsT<DfGeom, DfAccum>;

struct DfPgradArgs {
  DfGeom g;
  const float* packed;
  const float* xsave;        // [n_te][D][NL]
  const float* gsave;        // [n_te][D][NL]
  long n_te;
  int chunks_b;              // CTAs along the state-evaluation axis
  DfAccum acc;
};

struct DfPackArgs {
  DfGeom g;
  const float* Z;
  const float* ell;
  const float* var;
  const float* eps;
  const float* phase;
  const float* w;
  const float* nu;
  const float* B;
  float* packed;
};

struct DfFinalizeArgs {
  DfGeom g;
  const float* ell;
  const float* var;
  const float* w;
  DfAccum acc;
  float* d_Z;
  float* d_ell;
  float* d_var;
  float* d_nu;
  float* d_B;
};

DfGeom df_geom(const GpodeProblem* p);

cudaError_t df_launch_field_fwd(const DfFieldFwdArgs& a, cudaStream_t st);
cudaError_t df_launch_field_bwd(const DfFieldBwdArgs& a, cudaStream_t st);
cudaError_t df_launch_rollout_fwd(const DfRolloutFwdArgs& a, cudaStream_t st);
cudaError_t df_launch_rollout_bwd(const DfRolloutBwdArgs& a, cudaStream_t st);
cudaError_t df_launch_pgrad(const DfPgradArgs& a, cudaStream_t st);
cudaError_t df_launch_pack(const DfPackArgs& a, cudaStream_t st);
cudaError_t df_launch_finalize(const DfFinalizeArgs& a, cudaStream_t st);
constexpr int kDfSmallW = 8;   // warps per 32 states of the small-batch instantiation DfPolicy<D, 1, kDfSmallW>
// small batch: fewer states than one warp per SM can cover -- rows are split over the warps of a CTA and over a cluster instead
inline bool df_use_small(const DfGeom& g) { return static_cast<long>(g.N) * g.L <= 148L * 32; }
int df_smem_bytes(const DfGeom& g, int threads, int R, bool bwd, bool cluster = true, int W = 0);
int df_cluster(const DfGeom& g, int states_per_cta);

// per-D instantiations (df_inst.cu compiled once per GPODE_DF_D)
template <int D> cudaError_t df_field_fwd_d(const DfFieldFwdArgs& a, cudaStream_t st);
template <int D> cudaError_t df_field_bwd_d(const DfFieldBwdArgs& a, cudaStream_t st);
template <int D> cudaError_t df_rollout_fwd_d(const DfRolloutFwdArgs& a, cudaStream_t st);
template <int D> cudaError_t df_rollout_bwd_d(const DfRolloutBwdArgs& a, cudaStream_t st);
template <int D> cudaError_t df_pgrad_d(const DfPgradArgs& a, cudaStream_t st);

}  // namespace gpode
