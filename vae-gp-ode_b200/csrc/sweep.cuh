// Solver sweep kernels, generic over the field variant (policy P: RbfPolicy / DfPolicy).
//   k_field_fwd / k_field_bwd      one evaluation / one VJP, row-major I/O (SVGP_Layer.forward, torchdiffeq RHS)
//   k_rollout_fwd / k_rollout_bwd  whole fixed-grid solve / reverse sweep in ONE launch
// A policy supplies: Geom, Accum, Smem (with xs / dx staging buffers), DP, R, kThreads, kMinBlocks and
//   carve(), setup() -> total chunks, eval_fwd(store(k, fp, fu)), finish() (forward teardown), vjp() -> sum_k dx in sm.dx, flush().
#pragma once

#include "common.cuh"
#include "sweep_args.h"

namespace gpode {

// one thread's R states inside sample l
template <int R>
struct States {
  long s[R];    // global state index l*N + n (clamped for out-of-range lanes)
  bool ok[R];
};

// state_threads > 0: only the first `state_threads` threads of the CTA own states (the rest are helper warps of the policy)
template <int R, class G>
__device__ __forceinline__ States<R> map_states(const G& g, int state_threads = 0) {
  States<R> st;
  const int l = blockIdx.y;
  const int nthr = state_threads > 0 ? state_threads : static_cast<int>(blockDim.x);
  const int n0 = blockIdx.x * (nthr * R) + threadIdx.x;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int n = n0 + r * nthr;
    st.ok[r] = n < g.N && static_cast<int>(threadIdx.x) < nthr;
    st.s[r] = static_cast<long>(l) * g.N + (st.ok[r] ? n : g.N - 1);
  }
  return st;
}

// policies that fill global accumulators DURING the reverse sweep (the fused tcgen05 kernel) declare bind(Smem&, const Accum&)
template <class P, class S, class A>
__device__ __forceinline__ auto bind_accum(S& sm, const A& acc, int) -> decltype(P::bind(sm, acc), void()) { P::bind(sm, acc); }
template <class P, class S, class A>
__device__ __forceinline__ void bind_accum(S&, const A&, long) {}

// policies with helper warps may declare kCoopGlue = true (needs R == 1): the solver glue between two
// evaluations of a rollout -- loads of saves and adjoints, a dependent round trip to L2 / HBM per component when one thread
// walks the components of its state -- is then spread over ALL threads of the CTA, element (component d, state slot) <-> thread
// e % blockDim.x, every load of a thread independent (measured on the fused tcgen05 reverse sweep: 10 % of all warp time sat at the
// CTA barrier behind 128 state threads walking 16 components each)
template <class P, class = void>
struct coop_glue { static constexpr bool value = false; };
template <class P>
struct coop_glue<P, decltype((void)P::kCoopGlue, void())> { static constexpr bool value = P::kCoopGlue; };

// f(d, s, slot): component d of the state with global index s whose staging slot in the xs / dx buffers is `slot` (+ d * stride)
template <class P, class G, class F>
__device__ __forceinline__ void glue_each(const G& g, const States<P::R>& st, int DS, F&& f) {
  if constexpr (coop_glue<P>::value) {
    static_assert(P::R == 1 && P::kStateThreads > 0, "cooperative glue: one state per state thread");
    const int total = DS * P::kStateThreads;
    const int stride = P::kXsStride ? P::kXsStride : static_cast<int>(blockDim.x);
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      const int d = e / P::kStateThreads, sl = e - d * P::kStateThreads;
      const long n = static_cast<long>(blockIdx.x) * P::kStateThreads + sl;
      if (n < g.N) f(d, static_cast<long>(blockIdx.y) * g.N + n, d * stride + sl);
    }
  } else {
    constexpr int R = P::R;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (!st.ok[r]) continue;
#pragma unroll 1
      for (int d = 0; d < DS; ++d) f(d, st.s[r], (d * R + r) * (P::kXsStride ? P::kXsStride : static_cast<int>(blockDim.x)) + static_cast<int>(threadIdx.x));
    }
  }
}

#define GPODE_SWEEP_BOUNDS __launch_bounds__(P::kThreads, P::kMinBlocks)
#ifdef GPODE_BWD_MAXNREG   // timing experiments: register cap by __maxnreg__ instead of launch bounds
#define GPODE_SWEEP_BOUNDS_BWD __maxnreg__(GPODE_BWD_MAXNREG)
#else
#define GPODE_SWEEP_BOUNDS_BWD __launch_bounds__(P::kThreadsBwd, P::kMinBlocksBwd)
#endif

// smem slot of component d of this thread's r-th state (policies whose helper warps own no state may use a tighter stride: kXsStride)
#define GPODE_XS(buf, d, r) (buf)[((d) * R + (r)) * blockDim.x + threadIdx.x]
#define GPODE_XSG(buf, d, r) (buf)[((d) * R + (r)) * (P::kXsStride ? P::kXsStride : static_cast<int>(blockDim.x)) + threadIdx.x]

#define GPODE_POLICY_CONSTS constexpr int DP = P::DP; constexpr int R = P::R; (void)DP

// =============================================================================================
// field forward: one evaluation, row-major I/O
// =============================================================================================
template <class P>
__global__ void GPODE_SWEEP_BOUNDS k_field_fwd(const FieldFwdArgsT<typename P::Geom> a) {
  GPODE_POLICY_CONSTS;
  extern __shared__ __align__(128) float smem[];
  const typename P::Geom& g = a.g;
  const States<R> st = map_states<R>(g, P::kStateThreads);
  typename P::Smem sm = P::carve(smem, g);
  ChunkPipe pipe;
  const long total = P::setup(sm, pipe, g, a.packed, 1, false);
  if (P::kXsStride == 0 || static_cast<int>(threadIdx.x) < P::kStateThreads) {   // (helper warps of a tightly strided policy own no staging slot)
#pragma unroll
    for (int r = 0; r < R; ++r)
      for (int d = 0; d < g.D_in; ++d) GPODE_XSG(sm.xs, d, r) = a.x[st.s[r] * g.D_in + d];
  }
  P::eval_fwd(pipe, g, total, sm, [&](int k, const float (&fp)[R], const float (&fu)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (st.ok[r]) {
        a.f[st.s[r] * g.D_out + k] = fp[r] + fu[r];
        if (a.f_prior) a.f_prior[st.s[r] * g.D_out + k] = fp[r];
      }
  });
  P::finish(sm);
}

// =============================================================================================
// rollout forward: fixed-grid euler / midpoint / rk4(3/8) over ts, all stages, one launch.
// The state lives in global memory between evaluations (traj / save slabs written by this thread);
// only the current stage input is register resident.
// =============================================================================================
template <class P>
__global__ void GPODE_SWEEP_BOUNDS k_rollout_fwd(const RolloutFwdArgsT<typename P::Geom> a) {
  GPODE_POLICY_CONSTS;
  extern __shared__ __align__(128) float smem[];
  const typename P::Geom& g = a.g;
  const States<R> st = map_states<R>(g, P::kStateThreads);
  const int stages = a.method == GPODE_EULER ? 1 : (a.method == GPODE_MIDPOINT ? 2 : 4);
  const long NL = g.NL;
  const int DS = g.D_in;
  typename P::Smem sm = P::carve(smem, g);
  ChunkPipe pipe;
  const long total = P::setup(sm, pipe, g, a.packed, static_cast<long>(a.T - 1) * stages, false);

  glue_each<P>(g, st, DS, [&](int d, long s, int) {
    const long zrow = a.z0_per_sample ? s : (s - static_cast<long>(blockIdx.y) * g.N);
    a.traj[(s * a.T) * DS + d] = a.z0[zrow * DS + d];
  });
  float* ksave = a.ksave;  // plain pointers: values written below are re-read by the same thread
  float* traj = a.traj;
  const long ks = static_cast<long>(DS) * NL;  // K_j[d] sits j*ks after K_0[d]

#pragma unroll 1
  for (int t = 0; t < a.T - 1; ++t) {
    const float dt = a.ts[t + 1] - a.ts[t];
    const long slab = a.keep ? static_cast<long>(t) * stages : 0;
#pragma unroll 1
    for (int i = 0; i < stages; ++i) {
      // stage input, in the operation order of torchdiffeq's fixed-grid step functions (padded lanes keep the zeros of the staging buffer)
      glue_each<P>(g, st, DS, [&](int d, long s, int slot) {
        const float y0 = traj[(s * a.T + t) * DS + d];
        const long kb = (slab * DS + d) * NL + s;
        float v;
        if (i == 0) {
          v = y0;
        } else if (a.method == GPODE_MIDPOINT) {
          v = y0 + ksave[kb] * (0.5f * dt);
        } else if (i == 1) {
          v = y0 + dt * ksave[kb] * (1.f / 3.f);
        } else if (i == 2) {
          v = y0 + dt * (ksave[kb + ks] - ksave[kb] * (1.f / 3.f));
        } else {
          v = y0 + dt * (ksave[kb] - ksave[kb + ks] + ksave[kb + 2 * ks]);
        }
        a.xsave[((slab + i) * DS + d) * NL + s] = v;
        sm.xs[slot] = v;
        // order 2: the first q components of the derivative are the velocity part of the state
        if (g.order == 2 && d >= DP / 2) ksave[((slab + i) * DS + d - DP / 2) * NL + s] = v;
      });
      if constexpr (coop_glue<P>::value) __syncthreads();   // xs was staged by other threads than the ones that read it
      P::eval_fwd(pipe, g, total, sm, [&](int k, const float (&fp)[R], const float (&fu)[R]) {
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (st.ok[r]) {
            ksave[((slab + i) * DS + g.off + k) * NL + st.s[r]] = fp[r] + fu[r];
            a.fpsave[((slab + i) * g.D_out + k) * NL + st.s[r]] = fp[r];
          }
      });
      if constexpr (coop_glue<P>::value) __syncthreads();   // the stage derivatives (global) are read back by other threads
    }
    // step update
    glue_each<P>(g, st, DS, [&](int d, long s, int) {
      const float y0 = traj[(s * a.T + t) * DS + d];
      const long kb = (slab * DS + d) * NL + s;
      float v;
      if (a.method == GPODE_EULER) {
        v = y0 + dt * ksave[kb];
      } else if (a.method == GPODE_MIDPOINT) {
        v = y0 + dt * ksave[kb + ks];
      } else {
        v = y0 + (ksave[kb] + 3.f * (ksave[kb + ks] + ksave[kb + 2 * ks]) + ksave[kb + 3 * ks]) * dt * 0.125f;
      }
      traj[(s * a.T + (t + 1)) * DS + d] = v;
    });
  }
  P::finish(sm);
}

// =============================================================================================
// rollout backward: reverse sweep through the unrolled solver (stage adjoints via the tableau).
// Adjoint vectors live in global scratch ([component][state], coalesced).
// =============================================================================================
template <class P>
__global__ void GPODE_SWEEP_BOUNDS_BWD k_rollout_bwd(const RolloutBwdArgsT<typename P::Geom, typename P::Accum> a) {
  GPODE_POLICY_CONSTS;
  extern __shared__ __align__(128) float smem[];
  const typename P::Geom& g = a.g;
  const States<R> st = map_states<R>(g, P::kStateThreads);
  const Tableau tb = make_tableau(a.method);
  const int stages = tb.stages;
  const long NL = g.NL;
  const int DS = g.D_in;
  typename P::Smem sm = P::carve(smem, g);
  ChunkPipe pipe;
  const long total = P::setup(sm, pipe, g, a.packed, static_cast<long>(a.T - 1) * stages, true);
  bind_accum<P>(sm, a.acc, 0);
  float* ybar = a.ybar;
  float* ystage = a.ystage;
  float* kbar = a.kbar;

  // adjoint of z_{T-1}
  glue_each<P>(g, st, DS, [&](int d, long s, int) { ybar[d * NL + s] = a.dtraj[(s * a.T + (a.T - 1)) * DS + d]; });

#pragma unroll 1
  for (int t = a.T - 2; t >= 0; --t) {
    const float dt = a.ts[t + 1] - a.ts[t];
    const long slab = static_cast<long>(t) * stages;
#pragma unroll 1
    for (int i = stages - 1; i >= 0; --i) {
      // kbar_i = dt (b_i ybar + sum_{j>i} a_ji ybar_j); stage input back from the forward saves
      glue_each<P>(g, st, DS, [&](int d, long s, int slot) {
        float kb = tb.b[i] * ybar[d * NL + s];
        for (int j = i + 1; j < stages; ++j) kb = fmaf(tb.a[j][i], ystage[(j * DS + d) * NL + s], kb);
        kb *= dt;
        kbar[d * NL + s] = kb;
        if (d >= g.off) a.gsave[((slab + i) * g.D_out + (d - g.off)) * NL + s] = kb;
        sm.xs[slot] = a.xsave[((slab + i) * DS + d) * NL + s];
      });
      if constexpr (coop_glue<P>::value) __syncthreads();   // kbar (read by every warp of vjp) and xs were written by other threads
      P::vjp(pipe, g, total, sm, st, kbar + g.off * NL, a.ksave + ((slab + i) * DS + g.off) * NL, a.fpsave + (slab + i) * g.D_out * NL, NL,
             1);
      // ybar_i = J^T kbar_i (+ order 2: d(state derivative)[0:q] = state[q:2q], adjoint flows to the velocity part)
      glue_each<P>(g, st, DS, [&](int d, long s, int slot) {
        float v = sm.dx[slot];
        if (g.order == 2 && d >= DP / 2) v += kbar[(d - DP / 2) * NL + s];
        ystage[(i * DS + d) * NL + s] = v;
      });
      if constexpr (coop_glue<P>::value) __syncthreads();   // dx (aliases an operand buffer of the next evaluation) is consumed
    }
    // ybar_t = ybar_{t+1} + sum_i ybar_i + dL/dz_t
    glue_each<P>(g, st, DS, [&](int d, long s, int) {
      float v = ybar[d * NL + s] + a.dtraj[(s * a.T + t) * DS + d];
      for (int j = 0; j < stages; ++j) v += ystage[(j * DS + d) * NL + s];
      ybar[d * NL + s] = v;
    });
  }
  glue_each<P>(g, st, DS, [&](int d, long s, int) { a.dz0[s * DS + d] = ybar[d * NL + s]; });
  __syncthreads();
  P::flush(sm, g, a.acc);
}

// =============================================================================================
// field backward: one VJP, row-major I/O; leaves transposed x / g for the parameter-gradient kernel
// =============================================================================================
template <class P>
__global__ void GPODE_SWEEP_BOUNDS_BWD k_field_bwd(const FieldBwdArgsT<typename P::Geom, typename P::Accum> a) {
  GPODE_POLICY_CONSTS;
  extern __shared__ __align__(128) float smem[];
  const typename P::Geom& g = a.g;
  const States<R> st = map_states<R>(g, P::kStateThreads);
  const long NL = g.NL;
  typename P::Smem sm = P::carve(smem, g);
  ChunkPipe pipe;
  const long total = P::setup(sm, pipe, g, a.packed, 1, true);
  bind_accum<P>(sm, a.acc, 0);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (P::kXsStride == 0 || static_cast<int>(threadIdx.x) < P::kStateThreads)   // (helper warps of a tightly strided policy own no staging slot)
      for (int d = 0; d < g.D_in; ++d) {
        const float v = a.x[st.s[r] * g.D_in + d];
        GPODE_XSG(sm.xs, d, r) = v;
        if (st.ok[r]) a.xsave[d * NL + st.s[r]] = v;
      }
    if (st.ok[r])
      for (int k = 0; k < g.D_out; ++k) a.gsave[k * NL + st.s[r]] = a.gout[st.s[r] * g.D_out + k];
  }
  P::vjp(pipe, g, total, sm, st, a.gout, a.f, a.f_prior, 1, g.D_out);
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (st.ok[r])
      for (int d = 0; d < g.D_in; ++d) a.dx[st.s[r] * g.D_in + d] = GPODE_XSG(sm.dx, d, r);
  __syncthreads();
  P::flush(sm, g, a.acc);
}


}  // namespace gpode
