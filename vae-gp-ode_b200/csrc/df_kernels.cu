// Divergence-free kernel variant -- placeholder entry points until the DF kernels land.
#include "common.cuh"
#include "df.h"

namespace gpode {

DfGeom df_geom(const GpodeProblem* p) {
  DfGeom g;
  g.L = p->L; g.N = p->N; g.NL = p->L * p->N; g.D = p->D_in; g.DP = (p->D_in + 1) / 2 * 2;
  g.M = p->M; g.S = p->S; g.MP2 = (p->M + 1) / 2; g.SP2 = (p->S + 1) / 2;
  return g;
}
size_t df_packed_floats(const DfGeom&) { return 64; }
size_t df_acc_floats(const DfGeom&) { return 64; }
int df_field_fwd(const GpodeProblem*, const float*, float*, float*, float*, cudaStream_t) { return GPODE_E_UNSUPPORTED; }
int df_field_bwd(const GpodeProblem*, const float*, const float*, const float*, const float*, float*, const GpodeParamGrads*, float*,
                 float*, float*, float*, cudaStream_t) { return GPODE_E_UNSUPPORTED; }
int df_rollout_fwd(const GpodeProblem*, const float*, int, const float*, int, int, float*, float*, float*, float*, bool, float*,
                   cudaStream_t) { return GPODE_E_UNSUPPORTED; }
int df_rollout_bwd(const GpodeProblem*, const float*, int, int, const float*, const float*, const float*, const float*, float*,
                   const GpodeParamGrads*, float*, float*, float*, float*, float*, float*, cudaStream_t) { return GPODE_E_UNSUPPORTED; }

}  // namespace gpode
