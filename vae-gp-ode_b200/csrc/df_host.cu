// DF: geometry, parameter packing, gradient finalisation and the D dispatch of the sweep kernels.
#include "common.cuh"
#include "df.h"

namespace gpode {

DfGeom df_geom(const GpodeProblem* p) {
  DfGeom g;
  g.L = p->L;
  g.N = p->N;
  g.NL = p->L * p->N;
  g.D = g.D_in = g.D_out = p->D_in;
  g.M = p->M;
  g.S = p->S;
  g.MP2 = (p->M + 1) / 2;
  g.SP2 = (p->S + 1) / 2;
  g.rowf_s = (2 * g.D + 2) * 2;
  g.rowf_m = 2 * g.D * 2;
  g.hdr_floats = (4 * g.D * g.D + g.D + 3) / 4 * 4;
  const int rcs_max = kChunkBytes / (g.rowf_s * 4), rcm_max = kChunkBytes / (g.rowf_m * 4);
  g.NCs = (g.SP2 + rcs_max - 1) / rcs_max;
  g.RCs = (g.SP2 + g.NCs - 1) / g.NCs;
  g.NCm = (g.MP2 + rcm_max - 1) / rcm_max;
  g.RCm = (g.MP2 + g.NCm - 1) / g.NCm;
  const int fs = g.RCs * g.rowf_s, fm = g.RCm * g.rowf_m;
  g.stage_floats = fs > fm ? fs : fm;
  g.order = 1;
  g.off = 0;
  g.cg.stage_floats = g.stage_floats;
  g.cg.rowf_s = g.rowf_s;
  g.cg.rowf_m = g.rowf_m;
  g.cg.SP2 = g.SP2; g.cg.MP2 = g.MP2; g.cg.NCs = g.NCs; g.cg.NCm = g.NCm; g.cg.RCs = g.RCs; g.cg.RCm = g.RCm;
  g.cg.K = g.D;
  g.cg.blk_floats = g.SP2 * g.rowf_s;
  g.cg.tail = 1;
  return g;
}

namespace {
struct AccOff { size_t dnu, dz, dbp, dc, dell_x, dvar, total; };
AccOff acc_offsets(const DfGeom& g) {
  AccOff o;
  size_t at = 0;
  auto take = [&](size_t n) { const size_t a = at; at += (n + 63) / 64 * 64; return a; };
  o.dnu = take(static_cast<size_t>(g.L) * 2 * g.MP2 * g.D);
  o.dz = take(static_cast<size_t>(2) * g.MP2 * g.D);
  o.dbp = take(static_cast<size_t>(g.L) * g.D * 2 * g.SP2 * g.D);
  o.dc = take(g.D * g.D);
  o.dell_x = take(g.D * g.D);
  o.dvar = take(g.D);
  o.total = at;
  return o;
}
}  // namespace

size_t df_acc_floats(const DfGeom& g) { return acc_offsets(g).total; }
DfAccum df_acc(float* base, const DfGeom& g) {
  const AccOff o = acc_offsets(g);
  DfAccum a;
  a.dnu = base + o.dnu;
  a.dz = base + o.dz;
  a.dbp = base + o.dbp;
  a.dc = base + o.dc;
  a.dell_x = base + o.dell_x;
  a.dvar = base + o.dvar;
  return a;
}

// cluster size of a launch (rows of every chunk split over the CTAs of a cluster along grid z, df_kernels.cuh): small batches only --
// while the whole launch, cluster included, fits the chip once
// (states_per_cta: threads x R of the thread <-> state instantiations, 32 of the warp-split small-batch one)
int df_cluster(const DfGeom& g, int states_per_cta) {
  const long ctas = ((static_cast<long>(g.N) + states_per_cta - 1) / states_per_cta) * g.L;
  long c = 148 / (ctas > 0 ? ctas : 1);
  if (c > 8) c = 8;
  while (c > 1 && 2 * c * 2 * g.D * states_per_cta * 4 > 48 * 1024) --c;   // exchange buffers of at most 48 KB
  return c < 2 ? 1 : static_cast<int>(c);
}
// (cluster = false: the worst-case footprint without the exchange buffers of a cluster launch -- the shape check of the ABI;
//  W > 0: the warp-split instantiation, staging buffers of 32 states + the warps' partial-sum buffer)
int df_smem_bytes(const DfGeom& g, int threads, int R, bool bwd, bool cluster, int W) {
  const int states = W > 0 ? 32 : threads * R;
  int floats = 32 + kPipeStages * g.stage_floats + g.hdr_floats + g.D * states;
  const int C = cluster ? df_cluster(g, states) : 1;
  if (C > 1) floats += 2 * C * 2 * g.D * states;
  if (W > 0) floats += W * 2 * g.D * 32;
  if (bwd) floats += g.D * states + 2 * g.D * g.D + g.D + g.MP2 * 4 * g.D;
  return floats * 4;
}

// ---------------------------------------------------------------------------------------------
// pack (layout in df.h).  blockIdx.y = block a (feature rows) or D (inducing rows + header), blockIdx.z = sample
// ---------------------------------------------------------------------------------------------
__global__ void k_df_pack(const DfPackArgs a) {
  const DfGeom& g = a.g;
  const int D = g.D, S = g.S, M = g.M;
  const int blk = blockIdx.y, l = blockIdx.z;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  float* rows = const_cast<float*>(df_rows_ptr(a.packed, g, l));
  if (blk < D) {
    if (row >= g.SP2) return;
    const int fa = blk;
    float2* out = reinterpret_cast<float2*>(rows + (static_cast<size_t>(fa) * g.SP2 + row) * g.rowf_s);
    float om[kDfMaxD][2], bp[2] = {0.f, 0.f}, bb[kDfMaxD][2];
    for (int h = 0; h < 2; ++h) {
      const int s = 2 * row + h;
      const bool ok = s < S;
      float R = 0.f;
      if (ok) {
        const float w1 = a.w[(static_cast<size_t>(l) * 2 * S + s) * D + fa], w2 = a.w[(static_cast<size_t>(l) * 2 * S + S + s) * D + fa];
        R = hypotf(w1, w2);
        bp[h] = a.phase[(static_cast<size_t>(l) * S + s) * D + fa] - atan2f(w2, w1);
      }
      for (int d = 0; d < D; ++d) om[d][h] = ok ? a.eps[((static_cast<size_t>(l) * D + d) * S + s) * D + fa] / a.ell[fa * D + d] : 0.f;
      for (int c = 0; c < D; ++c)
        bb[c][h] = ok ? sqrtf(a.var[c] / static_cast<float>(S)) * R * a.B[((static_cast<size_t>(l) * S + s) * D + fa) * D + c] : 0.f;
    }
    for (int d = 0; d < D; ++d) out[d] = make_float2(om[d][0], om[d][1]);
    out[D] = make_float2(bp[0], bp[1]);
    for (int c = 0; c < D; ++c) out[D + 1 + c] = make_float2(bb[c][0], bb[c][1]);
    out[2 * D + 1] = make_float2(0.f, 0.f);
    return;
  }
  if (row < g.MP2) {
    float2* out = reinterpret_cast<float2*>(rows + static_cast<size_t>(D) * g.SP2 * g.rowf_s + static_cast<size_t>(row) * g.rowf_m);
    for (int d = 0; d < D; ++d) {
      float z[2] = {0.f, 0.f}, nu[2] = {0.f, 0.f};
      for (int h = 0; h < 2; ++h) {
        const int m = 2 * row + h;
        if (m < M) {
          z[h] = a.Z[m * D + d];
          nu[h] = a.nu[(static_cast<size_t>(l) * M + m) * D + d];
        }
      }
      out[d] = make_float2(z[0], z[1]);
      out[D + d] = make_float2(nu[0], nu[1]);
    }
  }
  if (l == 0 && row < D * D) {   // header: {k, k, lc, lc}_ij, then h_j
    const int i = row / D, j = row - i * D;
    const float e = a.ell[row], c = 1.f / (e * e);
    const float k = -0.5f * kLog2e * c, lc = log2f(a.var[j] * c * c);
    reinterpret_cast<float4*>(a.packed)[row] = make_float4(k, k, lc, lc);
    if (i == j) a.packed[4 * D * D + j] = static_cast<float>(D - 1) * e * e;
  }
}

cudaError_t df_launch_pack(const DfPackArgs& a, cudaStream_t st) {
  int rows = a.g.SP2 > a.g.MP2 ? a.g.SP2 : a.g.MP2;
  if (rows < a.g.D * a.g.D) rows = a.g.D * a.g.D;
  dim3 grid((rows + 127) / 128, a.g.D + 1, a.g.L);
  k_df_pack<<<grid, 128, 0, st>>>(a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// finalize: accumulators -> gradients in the reference layouts
//   d_nu (L,M*D,1), d_Z (M,D) copies;  d_ell_ij = -(dell_x_ij + 2 ln2 dc_ij) / ell_ij ;  d_var_j = dvar_j / var_j ;
//   d_B[l,s,a,c] = sqrt(var_c/S) R_sa dB'[l,a,s,c]
// ---------------------------------------------------------------------------------------------
__global__ void k_df_finalize(const DfFinalizeArgs a) {
  const DfGeom& g = a.g;
  const int D = g.D, M = g.M, S = g.S;
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  if (a.d_nu)
    for (long i = tid; i < static_cast<long>(g.L) * M * D; i += nthreads) {
      const long l = i / (static_cast<long>(M) * D), r = i - l * M * D;
      a.d_nu[i] = a.acc.dnu[l * 2 * g.MP2 * D + r];
    }
  if (a.d_Z)
    for (long i = tid; i < static_cast<long>(M) * D; i += nthreads) a.d_Z[i] = a.acc.dz[i];
  if (a.d_ell)
    for (long i = tid; i < D * D; i += nthreads) a.d_ell[i] = -(a.acc.dell_x[i] + 2.f * kLn2 * a.acc.dc[i]) / a.ell[i];
  if (a.d_var)
    for (long i = tid; i < D; i += nthreads) a.d_var[i] = a.acc.dvar[i] / a.var[i];
  if (a.d_B)
    for (long i = tid; i < static_cast<long>(g.L) * S * D * D; i += nthreads) {
      const int c = static_cast<int>(i % D);
      const int fa = static_cast<int>((i / D) % D);
      const int s = static_cast<int>((i / (D * D)) % S);
      const long l = i / (static_cast<long>(D) * D * S);
      const float w1 = a.w[(l * 2 * S + s) * D + fa], w2 = a.w[(l * 2 * S + S + s) * D + fa];
      const float scale = sqrtf(a.var[c] / static_cast<float>(S)) * hypotf(w1, w2);
      a.d_B[i] = scale * a.acc.dbp[((l * D + fa) * g.SP2 + s / 2) * 2 * D + (s & 1) * D + c];
    }
}

cudaError_t df_launch_finalize(const DfFinalizeArgs& a, cudaStream_t st) {
  k_df_finalize<<<64, 256, 0, st>>>(a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// D dispatch
// ---------------------------------------------------------------------------------------------
#define GPODE_DF_SWITCH(fn, a, st)            \
  switch ((a).g.D) {                          \
    case 1: return fn<1>(a, st);              \
    case 2: return fn<2>(a, st);              \
    case 3: return fn<3>(a, st);              \
    case 4: return fn<4>(a, st);              \
    case 5: return fn<5>(a, st);              \
    case 6: return fn<6>(a, st);              \
    case 7: return fn<7>(a, st);              \
    case 8: return fn<8>(a, st);              \
    default: return cudaErrorInvalidValue;    \
  }

cudaError_t df_launch_field_fwd(const DfFieldFwdArgs& a, cudaStream_t st) { GPODE_DF_SWITCH(df_field_fwd_d, a, st) }
cudaError_t df_launch_field_bwd(const DfFieldBwdArgs& a, cudaStream_t st) { GPODE_DF_SWITCH(df_field_bwd_d, a, st) }
cudaError_t df_launch_rollout_fwd(const DfRolloutFwdArgs& a, cudaStream_t st) { GPODE_DF_SWITCH(df_rollout_fwd_d, a, st) }
cudaError_t df_launch_rollout_bwd(const DfRolloutBwdArgs& a, cudaStream_t st) { GPODE_DF_SWITCH(df_rollout_bwd_d, a, st) }
cudaError_t df_launch_pgrad(const DfPgradArgs& a, cudaStream_t st) { GPODE_DF_SWITCH(df_pgrad_d, a, st) }

}  // namespace gpode
