// RBF (dimwise layout) sparse-GP vector field: fused forward / backward kernels for sm_100a.
//
// Math (reference experiments/model/core/kernels.py:64-110,140-153,174-181):
//   f_k(x) = sum_s w'_sk cos(x . omega_sk + b_sk)  +  sum_m nu'_km 2^( A_k(x) + x . G_km + H_km )
//   with w' = sqrt(var_k/S) w, nu' = var_k nu, c_kd = -log2(e)/(2 ell_kd^2), A_k = sum_d c_kd x_d^2,
//   G_kmd = -2 c_kd Z_md, H_km = sum_d c_kd Z_md^2   (A + x.G + H = sum_d c_kd (x_d - Z_md)^2).
// Mapping: R states per thread, the output dimension k is the OUTER loop so that only one
// parameter tile (omega_k, G_k ...) is live; tiles stream through shared memory (TilePipe).  Inside a
// tile, two features (or two inducing points) are processed per instruction with FFMA2: the state is
// the scalar-broadcast operand, the parameters come as {even,odd} pairs from broadcast LDS.128.
#pragma once

#include "common.cuh"
#include "rbf.h"

namespace gpode {

// one thread's R states inside sample l
template <int R>
struct States {
  long s[R];    // global state index l*N + n (clamped for out-of-range lanes)
  bool ok[R];
};

template <int R>
__device__ __forceinline__ States<R> map_states(const RbfGeom& g) {
  States<R> st;
  const int l = blockIdx.y;
  const int n0 = blockIdx.x * (blockDim.x * R) + threadIdx.x;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int n = n0 + r * blockDim.x;
    st.ok[r] = n < g.N;
    st.s[r] = static_cast<long>(l) * g.N + (st.ok[r] ? n : g.N - 1);
  }
  return st;
}

// ---------------------------------------------------------------------------------------------
// forward: prior part fp and update part fu of output k for R states
// ---------------------------------------------------------------------------------------------
template <int DP, int R>
__device__ __forceinline__ void rbf_tile_fwd(const float* __restrict__ tile, int SP2, int MP2, const float (&x)[R][DP],
                                             float (&fp)[R], float (&fu)[R]) {
  constexpr int HDR = rbf_hdr_floats(DP);
  constexpr int ROW4 = (DP + 2) / 2;  // float4 per row
  float A[R];
#pragma unroll
  for (int r = 0; r < R; ++r) A[r] = 0.f;
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    const float c = tile[d];
#pragma unroll
    for (int r = 0; r < R; ++r) A[r] = fmaf(c * x[r][d], x[r][d], A[r]);
  }
  const float4* rows = reinterpret_cast<const float4*>(tile + HDR);
  float2 acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll 2
  for (int j = 0; j < SP2; ++j) {
    float4 v[ROW4];
#pragma unroll
    for (int i = 0; i < ROW4; ++i) v[i] = rows[j * ROW4 + i];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 th = lo(v[ROW4 - 1]);
#pragma unroll
      for (int i = 0; i < DP / 2; ++i) {
        th = fma2(bc(x[r][2 * i]), lo(v[i]), th);
        th = fma2(bc(x[r][2 * i + 1]), hi(v[i]), th);
      }
      acc[r] = fma2(cos_2(th), hi(v[ROW4 - 1]), acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    fp[r] = acc[r].x + acc[r].y;
    acc[r] = make_float2(0.f, 0.f);
  }
  rows += static_cast<size_t>(SP2) * ROW4;
#pragma unroll 2
  for (int j = 0; j < MP2; ++j) {
    float4 v[ROW4];
#pragma unroll
    for (int i = 0; i < ROW4; ++i) v[i] = rows[j * ROW4 + i];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 ex = add2(lo(v[ROW4 - 1]), bc(A[r]));
#pragma unroll
      for (int i = 0; i < DP / 2; ++i) {
        ex = fma2(bc(x[r][2 * i]), lo(v[i]), ex);
        ex = fma2(bc(x[r][2 * i + 1]), hi(v[i]), ex);
      }
      acc[r] = fma2(ex2_2(ex), hi(v[ROW4 - 1]), acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) fu[r] = acc[r].x + acc[r].y;
}

// ---------------------------------------------------------------------------------------------
// backward (vector-Jacobian product) of output k: dxk[r][d] = g_k * d f_k / d x_d
//   rff:  -g sum_s w' sin(theta) omega_d ;  update: g ln2 (2 c_d x_d sum_m e_m + sum_m e_m G_md), e_m = nu'_m E_m
// ---------------------------------------------------------------------------------------------
template <int DP, int R>
__device__ __forceinline__ void rbf_tile_bwd(const float* __restrict__ tile, int SP2, int MP2, const float (&x)[R][DP],
                                             const float (&gk)[R], float (&dxk)[R][DP]) {
  constexpr int HDR = rbf_hdr_floats(DP);
  constexpr int ROW4 = (DP + 2) / 2;
  float A[R];
#pragma unroll
  for (int r = 0; r < R; ++r) A[r] = 0.f;
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    const float c = tile[d];
#pragma unroll
    for (int r = 0; r < R; ++r) A[r] = fmaf(c * x[r][d], x[r][d], A[r]);
  }
  const float4* rows = reinterpret_cast<const float4*>(tile + HDR);
  float2 Q[R][DP];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int d = 0; d < DP; ++d) Q[r][d] = make_float2(0.f, 0.f);
#pragma unroll 2
  for (int j = 0; j < SP2; ++j) {
    float4 v[ROW4];
#pragma unroll
    for (int i = 0; i < ROW4; ++i) v[i] = rows[j * ROW4 + i];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 th = lo(v[ROW4 - 1]);
#pragma unroll
      for (int i = 0; i < DP / 2; ++i) {
        th = fma2(bc(x[r][2 * i]), lo(v[i]), th);
        th = fma2(bc(x[r][2 * i + 1]), hi(v[i]), th);
      }
      const float2 t = mul2(hi(v[ROW4 - 1]), sin_2(th));
#pragma unroll
      for (int i = 0; i < DP / 2; ++i) {
        Q[r][2 * i] = fma2(t, lo(v[i]), Q[r][2 * i]);
        Q[r][2 * i + 1] = fma2(t, hi(v[i]), Q[r][2 * i + 1]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      dxk[r][d] = -gk[r] * (Q[r][d].x + Q[r][d].y);
      Q[r][d] = make_float2(0.f, 0.f);
    }
  rows += static_cast<size_t>(SP2) * ROW4;
  float2 Es[R];
#pragma unroll
  for (int r = 0; r < R; ++r) Es[r] = make_float2(0.f, 0.f);
#pragma unroll 2
  for (int j = 0; j < MP2; ++j) {
    float4 v[ROW4];
#pragma unroll
    for (int i = 0; i < ROW4; ++i) v[i] = rows[j * ROW4 + i];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 ex = add2(lo(v[ROW4 - 1]), bc(A[r]));
#pragma unroll
      for (int i = 0; i < DP / 2; ++i) {
        ex = fma2(bc(x[r][2 * i]), lo(v[i]), ex);
        ex = fma2(bc(x[r][2 * i + 1]), hi(v[i]), ex);
      }
      const float2 e = mul2(hi(v[ROW4 - 1]), ex2_2(ex));
      Es[r] = add2(Es[r], e);
#pragma unroll
      for (int i = 0; i < DP / 2; ++i) {
        Q[r][2 * i] = fma2(e, lo(v[i]), Q[r][2 * i]);
        Q[r][2 * i + 1] = fma2(e, hi(v[i]), Q[r][2 * i + 1]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const float es = Es[r].x + Es[r].y;
    const float gl = gk[r] * kLn2;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      const float c2 = 2.f * tile[d];
      dxk[r][d] = fmaf(gl, fmaf(c2 * x[r][d], es, Q[r][d].x + Q[r][d].y), dxk[r][d]);
    }
  }
}

// shared-memory carve-up of the sweep kernels: [2 mbarriers (pad to 128 B) | tile 0 | tile 1 | dell_x | dvar]
__device__ __forceinline__ float* smem_tiles(float* smem) { return smem + 32; }

// Folds this thread's contribution to the lengthscale (sum_n x_d dx_kd) and variance (sum_n g (f - fp/2))
// statistics of output k into the CTA accumulators: one warp reduction per value, lane 0 adds.
template <int DP, int R>
__device__ __forceinline__ void rbf_fold_stats(const float (&x)[R][DP], const float (&dxk)[R][DP], const float (&gk)[R],
                                               const float (&fk)[R], const float (&fpk)[R], const bool (&ok)[R], float* s_dell,
                                               float* s_dvar, int k) {
  const int lane = threadIdx.x & 31;
  float v = 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r) v += ok[r] ? gk[r] * (fk[r] - 0.5f * fpk[r]) : 0.f;
  v = warp_sum(v);
  if (lane == 0) atomicAdd(&s_dvar[k], v);
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    float u = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) u += ok[r] ? x[r][d] * dxk[r][d] : 0.f;
    u = warp_sum(u);
    if (lane == 0) atomicAdd(&s_dell[k * DP + d], u);
  }
}

// =============================================================================================
// field forward: one evaluation, row-major I/O
// =============================================================================================
template <int DP, int R>
__global__ void __launch_bounds__(256, 2) k_rbf_field_fwd(const RbfFieldFwdArgs a) {
  extern __shared__ __align__(128) float smem[];
  const RbfGeom& g = a.g;
  const States<R> st = map_states<R>(g);
  TilePipe pipe;
  pipe.init(smem_tiles(smem), reinterpret_cast<uint64_t*>(smem), a.packed + static_cast<size_t>(blockIdx.y) * g.D_out * g.tile_floats,
            g.tile_floats, g.D_out, g.D_out);
  float x[R][DP];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int d = 0; d < DP; ++d) x[r][d] = d < g.D_in ? a.x[st.s[r] * g.D_in + d] : 0.f;
  for (int k = 0; k < g.D_out; ++k) {
    const float* tile = pipe.acquire();
    float fp[R], fu[R];
    rbf_tile_fwd<DP, R>(tile, g.SP2, g.MP2, x, fp, fu);
    pipe.release();
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (st.ok[r]) {
        a.f[st.s[r] * g.D_out + k] = fp[r] + fu[r];
        if (a.f_prior) a.f_prior[st.s[r] * g.D_out + k] = fp[r];
      }
  }
}

// =============================================================================================
// rollout forward: fixed-grid euler / midpoint / rk4(3/8) over ts, all stages, one launch
// =============================================================================================
template <int DP, int R>
__global__ void __launch_bounds__(256, 2) k_rbf_rollout_fwd(const RbfRolloutFwdArgs a) {
  extern __shared__ __align__(128) float smem[];
  const RbfGeom& g = a.g;
  const States<R> st = map_states<R>(g);
  const int stages = a.method == GPODE_EULER ? 1 : (a.method == GPODE_MIDPOINT ? 2 : 4);
  const long NL = g.NL;
  const int DS = g.D_in;
  TilePipe pipe;
  pipe.init(smem_tiles(smem), reinterpret_cast<uint64_t*>(smem), a.packed + static_cast<size_t>(blockIdx.y) * g.D_out * g.tile_floats,
            g.tile_floats, g.D_out, static_cast<long>(a.T - 1) * stages * g.D_out);

  float y0[R][DP], x[R][DP];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long zrow = a.z0_per_sample ? st.s[r] : (st.s[r] - static_cast<long>(blockIdx.y) * g.N);
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      y0[r][d] = d < DS ? a.z0[zrow * DS + d] : 0.f;
      if (st.ok[r] && d < DS) a.traj[(st.s[r] * a.T) * DS + d] = y0[r][d];
    }
  }
  // K_j[d] of the current step, re-read from the save slab this thread wrote
  auto K = [&](long slab, int j, int d, int r) -> float { return a.ksave[((slab + j) * DS + d) * NL + st.s[r]]; };

#pragma unroll 1
  for (int t = 0; t < a.T - 1; ++t) {
    const float dt = a.ts[t + 1] - a.ts[t];
    const long slab = a.keep ? static_cast<long>(t) * stages : 0;
#pragma unroll 1
    for (int i = 0; i < stages; ++i) {
      // stage input (operation order of torchdiffeq's fixed-grid step functions)
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          float v = y0[r][d];
          if (d < DS && i > 0) {
            if (a.method == GPODE_MIDPOINT) {
              v = y0[r][d] + K(slab, 0, d, r) * (0.5f * dt);
            } else if (i == 1) {
              v = y0[r][d] + dt * K(slab, 0, d, r) * (1.f / 3.f);
            } else if (i == 2) {
              v = y0[r][d] + dt * (K(slab, 1, d, r) - K(slab, 0, d, r) * (1.f / 3.f));
            } else {
              v = y0[r][d] + dt * (K(slab, 0, d, r) - K(slab, 1, d, r) + K(slab, 2, d, r));
            }
          }
          x[r][d] = v;
          if (st.ok[r] && d < DS) a.xsave[((slab + i) * DS + d) * NL + st.s[r]] = v;
        }
      // order 2: the first q components of the derivative are the velocity part of the state
      if (g.order == 2) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int d = 0; d < DP / 2; ++d)
            if (st.ok[r]) a.ksave[((slab + i) * DS + d) * NL + st.s[r]] = x[r][d + DP / 2];
      }
      for (int k = 0; k < g.D_out; ++k) {
        const float* tile = pipe.acquire();
        float fp[R], fu[R];
        rbf_tile_fwd<DP, R>(tile, g.SP2, g.MP2, x, fp, fu);
        pipe.release();
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (st.ok[r]) {
            a.ksave[((slab + i) * DS + g.off + k) * NL + st.s[r]] = fp[r] + fu[r];
            a.fpsave[((slab + i) * g.D_out + k) * NL + st.s[r]] = fp[r];
          }
      }
    }
    // step update
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        if (d < DS) {
          float v;
          if (a.method == GPODE_EULER) {
            v = y0[r][d] + dt * K(slab, 0, d, r);
          } else if (a.method == GPODE_MIDPOINT) {
            v = y0[r][d] + dt * K(slab, 1, d, r);
          } else {
            v = y0[r][d] + (K(slab, 0, d, r) + 3.f * (K(slab, 1, d, r) + K(slab, 2, d, r)) + K(slab, 3, d, r)) * dt * 0.125f;
          }
          y0[r][d] = v;
          if (st.ok[r]) a.traj[(st.s[r] * a.T + (t + 1)) * DS + d] = v;
        }
      }
  }
}

// =============================================================================================
// rollout backward: reverse sweep through the unrolled solver (stage adjoints via the tableau)
// =============================================================================================
template <int DP, int R>
__global__ void __launch_bounds__(256, (DP <= 8 ? 2 : 1)) k_rbf_rollout_bwd(const RbfRolloutBwdArgs a) {
  extern __shared__ __align__(128) float smem[];
  const RbfGeom& g = a.g;
  const States<R> st = map_states<R>(g);
  const Tableau tb = make_tableau(a.method);
  const int stages = tb.stages;
  const long NL = g.NL;
  const int DS = g.D_in;
  float* s_dell = smem_tiles(smem) + 2 * g.tile_floats;
  float* s_dvar = s_dell + g.D_out * DP;
  for (int i = threadIdx.x; i < g.D_out * (DP + 1); i += blockDim.x) s_dell[i] = 0.f;
  TilePipe pipe;  // init() contains the __syncthreads that publishes the zeroing above
  pipe.init(smem_tiles(smem), reinterpret_cast<uint64_t*>(smem), a.packed + static_cast<size_t>(blockIdx.y) * g.D_out * g.tile_floats,
            g.tile_floats, g.D_out, static_cast<long>(a.T - 1) * stages * g.D_out);

  // adjoint of z_{T-1}
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int d = 0; d < DP; ++d)
      if (d < DS) a.ybar[d * NL + st.s[r]] = st.ok[r] ? a.dtraj[(st.s[r] * a.T + (a.T - 1)) * DS + d] : 0.f;

#pragma unroll 1
  for (int t = a.T - 2; t >= 0; --t) {
    const float dt = a.ts[t + 1] - a.ts[t];
    const long slab = static_cast<long>(t) * stages;
#pragma unroll 1
    for (int i = stages - 1; i >= 0; --i) {
      float x[R][DP], dx[R][DP];
      // kbar_i = dt (b_i ybar + sum_{j>i} a_ji ybar_j)
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          dx[r][d] = 0.f;
          x[r][d] = 0.f;
          if (d < DS) {
            float kb = tb.b[i] * a.ybar[d * NL + st.s[r]];
            for (int j = i + 1; j < stages; ++j) kb = fmaf(tb.a[j][i], a.ystage[(j * DS + d) * NL + st.s[r]], kb);
            kb *= dt;
            if (st.ok[r]) {
              a.kbar[d * NL + st.s[r]] = kb;
              if (d >= g.off) a.gsave[((slab + i) * g.D_out + (d - g.off)) * NL + st.s[r]] = kb;
            }
            x[r][d] = a.xsave[((slab + i) * DS + d) * NL + st.s[r]];
          }
        }
      for (int k = 0; k < g.D_out; ++k) {
        float gk[R], fk[R], fpk[R], dxk[R][DP];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          gk[r] = st.ok[r] ? a.kbar[(g.off + k) * NL + st.s[r]] : 0.f;
          fk[r] = a.ksave[((slab + i) * DS + g.off + k) * NL + st.s[r]];
          fpk[r] = a.fpsave[((slab + i) * g.D_out + k) * NL + st.s[r]];
        }
        const float* tile = pipe.acquire();
        rbf_tile_bwd<DP, R>(tile, g.SP2, g.MP2, x, gk, dxk);
        pipe.release();
        rbf_fold_stats<DP, R>(x, dxk, gk, fk, fpk, st.ok, s_dell, s_dvar, k);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int d = 0; d < DP; ++d) dx[r][d] += dxk[r][d];
      }
      // order 2: d(state derivative)[0:q] = state[q:2q]  ->  adjoint flows straight to the velocity part
      if (g.order == 2) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int d = 0; d < DP / 2; ++d) dx[r][d + DP / 2] += a.kbar[d * NL + st.s[r]];
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int d = 0; d < DP; ++d)
          if (d < DS && st.ok[r]) a.ystage[(i * DS + d) * NL + st.s[r]] = dx[r][d];
    }
    // ybar_t = ybar_{t+1} + sum_i ybar_i + dL/dz_t
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int d = 0; d < DP; ++d)
        if (d < DS && st.ok[r]) {
          float v = a.ybar[d * NL + st.s[r]] + a.dtraj[(st.s[r] * a.T + t) * DS + d];
          for (int j = 0; j < stages; ++j) v += a.ystage[(j * DS + d) * NL + st.s[r]];
          a.ybar[d * NL + st.s[r]] = v;
        }
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int d = 0; d < DP; ++d)
      if (d < DS && st.ok[r]) a.dz0[st.s[r] * DS + d] = a.ybar[d * NL + st.s[r]];
  __syncthreads();
  for (int i = threadIdx.x; i < g.D_out * DP; i += blockDim.x) atomicAdd(&a.acc.dell_x[i], s_dell[i]);
  for (int i = threadIdx.x; i < g.D_out; i += blockDim.x) atomicAdd(&a.acc.dvar[i], s_dvar[i]);
}

// =============================================================================================
// field backward: one VJP, row-major I/O; leaves transposed x / g for the parameter-gradient kernel
// =============================================================================================
template <int DP, int R>
__global__ void __launch_bounds__(256, (DP <= 8 ? 2 : 1)) k_rbf_field_bwd(const RbfFieldBwdArgs a) {
  extern __shared__ __align__(128) float smem[];
  const RbfGeom& g = a.g;
  const States<R> st = map_states<R>(g);
  const long NL = g.NL;
  float* s_dell = smem_tiles(smem) + 2 * g.tile_floats;
  float* s_dvar = s_dell + g.D_out * DP;
  for (int i = threadIdx.x; i < g.D_out * (DP + 1); i += blockDim.x) s_dell[i] = 0.f;
  TilePipe pipe;
  pipe.init(smem_tiles(smem), reinterpret_cast<uint64_t*>(smem), a.packed + static_cast<size_t>(blockIdx.y) * g.D_out * g.tile_floats,
            g.tile_floats, g.D_out, g.D_out);
  float x[R][DP], dx[R][DP];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      x[r][d] = d < g.D_in ? a.x[st.s[r] * g.D_in + d] : 0.f;
      dx[r][d] = 0.f;
      if (d < g.D_in && st.ok[r]) a.xsave[d * NL + st.s[r]] = x[r][d];
    }
  for (int k = 0; k < g.D_out; ++k) {
    float gk[R], fk[R], fpk[R], dxk[R][DP];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      gk[r] = st.ok[r] ? a.gout[st.s[r] * g.D_out + k] : 0.f;
      fk[r] = a.f[st.s[r] * g.D_out + k];
      fpk[r] = a.f_prior[st.s[r] * g.D_out + k];
      if (st.ok[r]) a.gsave[k * NL + st.s[r]] = gk[r];
    }
    const float* tile = pipe.acquire();
    rbf_tile_bwd<DP, R>(tile, g.SP2, g.MP2, x, gk, dxk);
    pipe.release();
    rbf_fold_stats<DP, R>(x, dxk, gk, fk, fpk, st.ok, s_dell, s_dvar, k);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int d = 0; d < DP; ++d) dx[r][d] += dxk[r][d];
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int d = 0; d < DP; ++d)
      if (d < g.D_in && st.ok[r]) a.dx[st.s[r] * g.D_in + d] = dx[r][d];
  __syncthreads();
  for (int i = threadIdx.x; i < g.D_out * DP; i += blockDim.x) atomicAdd(&a.acc.dell_x[i], s_dell[i]);
  for (int i = threadIdx.x; i < g.D_out; i += blockDim.x) atomicAdd(&a.acc.dvar[i], s_dvar[i]);
}

// =============================================================================================
// parameter gradients: threads <-> inducing-point pairs of one (sample, output dim); the CTA walks a
// chunk of state evaluations staged through shared memory and accumulates in registers:
//   dnu'_m = sum_n g_n E_nm ,  pg_md = sum_n g_n E_nm x_nd
// =============================================================================================
constexpr int kPgBatch = 128;

template <int DP>
__global__ void __launch_bounds__(256) k_rbf_pgrad(const RbfPgradArgs a) {
  const RbfGeom& g = a.g;
  constexpr int HDR = rbf_hdr_floats(DP);
  constexpr int ROW4 = (DP + 2) / 2;
  constexpr int SROW = ((DP + 2 + 3) / 4) * 4;  // staged state: x[DP], g, A (+pad), 16-byte rows
  constexpr int NV = SROW / 4;
  __shared__ __align__(16) float stage[kPgBatch * SROW];
  __shared__ float s_c[DP];
  const int k = blockIdx.y, l = blockIdx.z;
  const float* tile = a.packed + (static_cast<size_t>(l) * g.D_out + k) * g.tile_floats;
  if (threadIdx.x < DP) s_c[threadIdx.x] = tile[threadIdx.x];
  const int j = threadIdx.x;  // inducing pair
  const bool active = j < g.MP2;
  float2 G[DP], H = make_float2(0.f, 0.f);
#pragma unroll
  for (int d = 0; d < DP; ++d) G[d] = make_float2(0.f, 0.f);
  if (active) {
    const float4* row = reinterpret_cast<const float4*>(tile + HDR) + static_cast<size_t>(g.SP2 + j) * ROW4;
#pragma unroll
    for (int i = 0; i < DP / 2; ++i) {
      const float4 v = row[i];
      G[2 * i] = lo(v);
      G[2 * i + 1] = hi(v);
    }
    H = lo(row[ROW4 - 1]);
  }
  float2 dnu = make_float2(0.f, 0.f), pg[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) pg[d] = make_float2(0.f, 0.f);

  const long total = a.n_te * g.N;  // state evaluations of this sample
  const long per = (total + a.chunks - 1) / a.chunks;
  const long e_lo = static_cast<long>(blockIdx.x) * per;
  const long e_hi = e_lo + per < total ? e_lo + per : total;
  __syncthreads();
  for (long e0 = e_lo; e0 < e_hi; e0 += kPgBatch) {
    for (int idx = threadIdx.x; idx < kPgBatch; idx += blockDim.x) {
      const long e = e0 + idx;
      float xs[DP], gg = 0.f, A = 0.f;
#pragma unroll
      for (int d = 0; d < DP; ++d) xs[d] = 0.f;
      if (e < e_hi) {
        const long te = e / g.N;
        const long s = static_cast<long>(l) * g.N + (e - te * g.N);
#pragma unroll
        for (int d = 0; d < DP; ++d)
          if (d < g.D_in) xs[d] = a.xsave[(te * g.D_in + d) * g.NL + s];
        gg = a.gsave[(te * g.D_out + k) * g.NL + s];
#pragma unroll
        for (int d = 0; d < DP; ++d) A = fmaf(s_c[d] * xs[d], xs[d], A);
      }
      float* row = stage + idx * SROW;
#pragma unroll
      for (int d = 0; d < DP; ++d) row[d] = xs[d];
      row[DP] = gg;
      row[DP + 1] = A;
    }
    __syncthreads();
    if (active) {
      const int nb = (e_hi - e0) < kPgBatch ? static_cast<int>(e_hi - e0) : kPgBatch;
#pragma unroll 2
      for (int idx = 0; idx < nb; ++idx) {
        const float4* row = reinterpret_cast<const float4*>(stage + idx * SROW);
        float xv[SROW];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float4 q = row[i];
          xv[4 * i] = q.x;
          xv[4 * i + 1] = q.y;
          xv[4 * i + 2] = q.z;
          xv[4 * i + 3] = q.w;
        }
        float2 ex = add2(H, bc(xv[DP + 1]));
#pragma unroll
        for (int d = 0; d < DP; ++d) ex = fma2(bc(xv[d]), G[d], ex);
        const float2 ge = mul2(bc(xv[DP]), ex2_2(ex));
        dnu = add2(dnu, ge);
#pragma unroll
        for (int d = 0; d < DP; ++d) pg[d] = fma2(bc(xv[d]), ge, pg[d]);
      }
    }
    __syncthreads();
  }
  if (active) {
    const size_t base = (static_cast<size_t>(l) * g.D_out + k) * (2 * g.MP2) + 2 * j;
    atomicAdd(&a.acc.dnu[base], dnu.x);
    atomicAdd(&a.acc.dnu[base + 1], dnu.y);
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      atomicAdd(&a.acc.pg[base * DP + d], pg[d].x);
      atomicAdd(&a.acc.pg[(base + 1) * DP + d], pg[d].y);
    }
  }
}

// launch-shape heuristic of the sweep kernels: states per CTA = threads * R
inline void rbf_pick_shape(const RbfGeom& g, int& threads, int& R) {
  const long states = static_cast<long>(g.N);
  const long want = 2L * 148;  // CTAs over all samples for a full chip
  const int cand[5][2] = {{256, 2}, {256, 1}, {128, 1}, {64, 1}, {32, 1}};
  for (int i = 0; i < 5; ++i) {
    threads = cand[i][0];
    R = cand[i][1];
    if (R == 2 && g.DP > 8) continue;
    const long per = static_cast<long>(threads) * R;
    if (((states + per - 1) / per) * g.L >= want) return;
  }
}

}  // namespace gpode
