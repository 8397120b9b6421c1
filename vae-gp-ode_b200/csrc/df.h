// Divergence-free kernel variant: host-side entry points used by the C ABI (df_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "gpode.h"

namespace gpode {

constexpr int kDfMaxD = 8;

struct DfGeom {
  int L, N, NL, D, DP, M, S, MP2, SP2;
};

DfGeom df_geom(const GpodeProblem* p);
size_t df_packed_floats(const DfGeom& g);
size_t df_acc_floats(const DfGeom& g);

int df_field_fwd(const GpodeProblem* p, const float* x, float* f, float* f_prior, float* packed, cudaStream_t st);
int df_field_bwd(const GpodeProblem* p, const float* x, const float* g_out, const float* f, const float* f_prior, float* dx,
                 const GpodeParamGrads* grads, float* packed, float* xsave, float* gsave, float* acc, cudaStream_t st);
int df_rollout_fwd(const GpodeProblem* p, const float* z0, int z0_per_sample, const float* ts, int T, int method, float* traj,
                   float* xs, float* ks, float* fps, bool keep, float* packed, cudaStream_t st);
int df_rollout_bwd(const GpodeProblem* p, const float* ts, int T, int method, const float* xs, const float* ks, const float* fps,
                   const float* dtraj, float* dz0, const GpodeParamGrads* grads, float* packed, float* gsave, float* ybar,
                   float* ystage, float* kbar, float* acc, cudaStream_t st);

}  // namespace gpode
