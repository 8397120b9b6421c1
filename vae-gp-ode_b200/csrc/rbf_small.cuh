// RBF sweep policy for SMALL batches (the shape the reference actually trains at: 25 .. 256 trajectories, BASELINE config 1 / 3).
//
// With a few dozen states the regular kernels are latency bound: one warp owns 32 states and walks all D_out (S + M) parameter rows
// serially (~100 us per field evaluation forward + backward at config 1).  Here a CTA still owns 32 states (lane <-> state) but
// brings 16 warps: the parameter rows of the sample are loaded into shared memory ONCE per launch (config 1: 68 KB -- no streaming
// pipeline, no barrier per chunk), every warp evaluates an interleaved 1/16 of the rows of output k for all 32 states, and the
// partial sums meet through a double-buffered shared-memory exchange (one CTA barrier per output dimension).  Warp 0 -- the state
// threads of the generic solver glue (sweep.cuh) -- finishes each output: prior / update parts in the forward, dx_k and the
// lengthscale / variance statistics in the reverse sweep.  Arithmetic per row is the FFMA family's (row_dot / row_bwd_one).
#pragma once

#include "rbf_kernels.cuh"

namespace gpode {

constexpr int kSmWarps = 16;
constexpr int kSmThreads = kSmWarps * 32;
constexpr int kSmStates = 32;

struct SmallSmem {
  float* rows;   // [D_out][SP2 + MP2][ROWF] of this sample
  float* hdr;    // [D_out][hdr_floats]
  float* xs;     // [DP][32]
  float* dx;     // [DP][32]
  float* red;    // [2][kSmWarps][DP + 2][32] partial sums
  float* dell;   // [D_out][DP], then dvar [D_out]
  float* dvar;
};

inline int rbf_small_smem_bytes(const RbfGeom& g) {
  return (g.D_out * (g.SP2 + g.MP2) * g.row_floats + g.D_out * g.hdr_floats + 2 * g.DP * kSmStates + 2 * kSmWarps * (g.DP + 2) * kSmStates +
          g.D_out * (g.DP + 1) + 8) * 4;
}
// small batch and a parameter set that fits in shared memory next to the exchange buffers
inline bool rbf_use_small(const RbfGeom& g) {
  return static_cast<long>(g.N) * g.L <= 148L * 32 && rbf_small_smem_bytes(g) <= 200 * 1024;
}

template <int DP_>
struct RbfSmallPolicy {
  static constexpr int DP = DP_;
  static constexpr int R = 1;
  static constexpr int kThreads = kSmThreads;
  static constexpr int kMinBlocks = 1;
  static constexpr int kStateThreads = kSmStates;
  static constexpr int kXsStride = kSmStates;
  static constexpr bool kCoopGlue = true;   // sweep.cuh: the solver glue (dependent L2 round trips) is spread over the CTA's 512 threads
  static constexpr int kThreadsBwd = kSmThreads;
  static constexpr int kMinBlocksBwd = 1;
  using Geom = RbfGeom;
  using Accum = RbfAccum;
  using Smem = SmallSmem;
  static constexpr int ROW4 = (DP_ + 2) / 2;
  static constexpr int RED = DP_ + 2;

  __device__ static __forceinline__ Smem carve(float* smem, const Geom& g) {
    Smem s;
    s.rows = smem;
    s.hdr = s.rows + g.D_out * (g.SP2 + g.MP2) * g.row_floats;
    s.xs = s.hdr + g.D_out * g.hdr_floats;
    s.dx = s.xs + DP * kSmStates;
    s.red = s.dx + DP * kSmStates;
    s.dell = s.red + 2 * kSmWarps * RED * kSmStates;
    s.dvar = s.dell + g.D_out * DP;
    return s;
  }
  __device__ static __forceinline__ long setup(Smem& sm, ChunkPipe&, const Geom& g, const float* packed, long, bool) {
    const int l = blockIdx.y;
    const float4* src = reinterpret_cast<const float4*>(rbf_rows_ptr(packed, g, l));
    float4* dst = reinterpret_cast<float4*>(sm.rows);
    const int n4 = g.D_out * (g.SP2 + g.MP2) * g.row_floats / 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
    const float* hdr = rbf_hdr_ptr(packed, g, l);
    for (int i = threadIdx.x; i < g.D_out * g.hdr_floats; i += blockDim.x) sm.hdr[i] = hdr[i];
    for (int i = threadIdx.x; i < 2 * DP * kSmStates; i += blockDim.x) sm.xs[i] = 0.f;   // xs and dx (padded components stay 0)
    for (int i = threadIdx.x; i < g.D_out * (DP + 1); i += blockDim.x) sm.dell[i] = 0.f;  // dell and dvar
    __syncthreads();
    return 0;
  }
  __device__ static __forceinline__ void finish(Smem&) {}
  __device__ static __forceinline__ void flush(const Smem& sm, const Geom& g, const Accum& acc) {
    for (int i = threadIdx.x; i < g.D_out * DP; i += blockDim.x) atomicAdd(&acc.dell_x[i], sm.dell[i]);
    for (int i = threadIdx.x; i < g.D_out; i += blockDim.x) atomicAdd(&acc.dvar[i], sm.dvar[i]);
  }

  template <class Store>
  __device__ static __forceinline__ void eval_fwd(ChunkPipe&, const Geom& g, long, const Smem& sm, Store&& store) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();   // the stage input staged by warp 0 is visible; the previous evaluation's exchange buffers are free
    float x[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) x[d] = sm.xs[d * kSmStates + lane];
    const int nrows = g.SP2 + g.MP2;
    for (int k = 0; k < g.D_out; ++k) {
      const float* hdr_k = sm.hdr + k * g.hdr_floats;
      float A = 0.f;
#pragma unroll
      for (int d = 0; d < DP; ++d) A = fmaf(hdr_k[d] * x[d], x[d], A);
      const float4* rows = reinterpret_cast<const float4*>(sm.rows + static_cast<size_t>(k) * nrows * g.row_floats);
      float2 accp = make_float2(0.f, 0.f), accu = make_float2(0.f, 0.f);
      for (int j = warp; j < g.SP2; j += kSmWarps) {
        float4 a[ROW4];
        load_row<DP>(a, rows + j * ROW4);
        accp = fma2(cos_2(row_dot<DP, 1>(a, x, 0.f)), hi(a[ROW4 - 1]), accp);
      }
      for (int j = g.SP2 + warp; j < nrows; j += kSmWarps) {
        float4 a[ROW4];
        load_row<DP>(a, rows + j * ROW4);
        accu = fma2(ex2_2(row_dot<DP, 1>(a, x, A)), hi(a[ROW4 - 1]), accu);
      }
      float* red = sm.red + ((k & 1) * kSmWarps + warp) * RED * kSmStates;
      red[lane] = accp.x + accp.y;
      red[kSmStates + lane] = accu.x + accu.y;
      __syncthreads();
      if (warp == 0) {
        float fp[1] = {0.f}, fu[1] = {0.f};
        const float* all = sm.red + (k & 1) * kSmWarps * RED * kSmStates;
#pragma unroll
        for (int w = 0; w < kSmWarps; ++w) {
          fp[0] += all[w * RED * kSmStates + lane];
          fu[0] += all[w * RED * kSmStates + kSmStates + lane];
        }
        fu[0] *= kInvLn2;   // inducing-row weights carry ln2
        store(k, fp, fu);
      }
    }
  }

  __device__ static __forceinline__ void vjp(ChunkPipe&, const Geom& g, long, const Smem& sm, const States<R>& st, const float* gvec, const float* fvec,
                                             const float* fpvec, long kstride, long sstride) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    float x[1][DP], dxs[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      x[0][d] = sm.xs[d * kSmStates + lane];
      dxs[d] = 0.f;
    }
    const int nrows = g.SP2 + g.MP2;
    for (int k = 0; k < g.D_out; ++k) {
      const float* hdr_k = sm.hdr + k * g.hdr_floats;
      float A = 0.f, Q[DP], Es = 0.f;
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        A = fmaf(hdr_k[d] * x[0][d], x[0][d], A);
        Q[d] = 0.f;
      }
      const float4* rows = reinterpret_cast<const float4*>(sm.rows + static_cast<size_t>(k) * nrows * g.row_floats);
      for (int j = warp; j < g.SP2; j += kSmWarps) {
        float4 a[ROW4];
        load_row<DP>(a, rows + j * ROW4);
        row_bwd_one<DP, 1, false>(a, hi(a[ROW4 - 1]), x[0], kHalfPi, Q, Es);
      }
      for (int j = g.SP2 + warp; j < nrows; j += kSmWarps) {
        float4 a[ROW4];
        load_row<DP>(a, rows + j * ROW4);
        row_bwd_one<DP, 1, true>(a, hi(a[ROW4 - 1]), x[0], A, Q, Es);
      }
      float* red = sm.red + ((k & 1) * kSmWarps + warp) * RED * kSmStates;
#pragma unroll
      for (int d = 0; d < DP; ++d) red[d * kSmStates + lane] = Q[d];
      red[DP * kSmStates + lane] = Es;
      __syncthreads();
      if (warp == 0) {
        const float* all = sm.red + (k & 1) * kSmWarps * RED * kSmStates;
        float es = 0.f, dxk[1][DP];
#pragma unroll
        for (int d = 0; d < DP; ++d) dxk[0][d] = 0.f;
#pragma unroll 4
        for (int w = 0; w < kSmWarps; ++w) {
#pragma unroll
          for (int d = 0; d < DP; ++d) dxk[0][d] += all[(w * RED + d) * kSmStates + lane];
          es += all[(w * RED + DP) * kSmStates + lane];
        }
        const long at = k * kstride + st.s[0] * sstride;
        float gk[1], fk[1], fpk[1];
        gk[0] = st.ok[0] ? gvec[at] : 0.f;
        fk[0] = fvec[at];
        fpk[0] = fpvec[at];
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          dxk[0][d] = gk[0] * fmaf(2.f * hdr_k[d] * x[0][d], es, dxk[0][d]);
          dxs[d] += dxk[0][d];
        }
        rbf_fold_stats<DP, 1>(x, dxk, gk, fk, fpk, st.ok, sm.dell, sm.dvar, k);
      }
    }
    if (warp == 0) {
#pragma unroll
      for (int d = 0; d < DP; ++d) sm.dx[d * kSmStates + lane] = dxs[d];
    }
    __syncthreads();   // dx is consumed by the cooperative glue (all threads)
  }
};

}  // namespace gpode
