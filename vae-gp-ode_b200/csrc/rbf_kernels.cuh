// RBF (dimwise layout) sparse-GP vector field: fused forward / backward kernels for sm_100a.
//
// Math (reference experiments/model/core/kernels.py:64-110,140-153,174-181):
//   f_k(x) = sum_s w'_sk cos(x . omega_sk + b_sk)  +  sum_m nu'_km 2^( A_k(x) + x . G_km + H_km )
//   with w' = sqrt(var_k/S) w, nu' = var_k nu, c_kd = -log2(e)/(2 ell_kd^2), A_k = sum_d c_kd x_d^2,
//   G_kmd = -2 c_kd Z_md, H_km = sum_d c_kd Z_md^2   (A + x.G + H = sum_d c_kd (x_d - Z_md)^2).
//
// Mapping: R states per thread (register blocking: one broadcast LDS.128 of parameters feeds R states;
// with R = 1 the kernel is LDS-bound), output dimension k is the OUTER loop so that only
// the rows of one k are live; rows stream through shared memory in chunks (ChunkPipe, 1-D TMA).  Inside
// a chunk two features (or two inducing points) are processed per instruction with FFMA2: the state is
// the scalar-broadcast operand, the parameters come as {even, odd} pairs.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "rbf.h"
#include "sweep.cuh"

namespace gpode {

constexpr float kHalfPi = 1.5707963267948966f;
constexpr float kInvLn2 = 1.4426950408889634f;

template <int DP>
__device__ __forceinline__ void load_row(float4 (&v)[(DP + 2) / 2], const float4* row) {
#pragma unroll
  for (int i = 0; i < (DP + 2) / 2; ++i) v[i] = row[i];
}

// theta pair (two features / inducing points) of one state: row offset term + add + x . row
template <int DP, int R>
__device__ __forceinline__ float2 row_dot(const float4 (&v)[(DP + 2) / 2], const float (&x)[DP], float add) {
  constexpr int ROW4 = (DP + 2) / 2;
  float2 t0 = add2(lo(v[ROW4 - 1]), bc(add));
  if constexpr (R == 1 && DP >= 4) {  // single state per thread: split the dependent chain in two (measured: hurts at R = 2)
    float2 t1 = mul2(bc(x[1]), hi(v[0]));
    t0 = fma2(bc(x[0]), lo(v[0]), t0);
#pragma unroll
    for (int i = 1; i < DP / 2; ++i) {
      t0 = fma2(bc(x[2 * i]), lo(v[i]), t0);
      t1 = fma2(bc(x[2 * i + 1]), hi(v[i]), t1);
    }
    return add2(t0, t1);
  } else {
#pragma unroll
    for (int i = 0; i < DP / 2; ++i) {
      t0 = fma2(bc(x[2 * i]), lo(v[i]), t0);
      t0 = fma2(bc(x[2 * i + 1]), hi(v[i]), t0);
    }
    return t0;
  }
}

// ---------------------------------------------------------------------------------------------
// forward over one chunk of n rows: acc += weight * (IS_K ? 2^(theta) : cos(theta)).  Two rows per
// iteration, all their LDS.128 issued first: 2R independent FFMA2 chains per thread.
// ---------------------------------------------------------------------------------------------
template <int DP, int R, bool IS_K>
__device__ __forceinline__ void rows_fwd(const float* __restrict__ chunk, int n, const float (&x)[R][DP], const float (&A)[R],
                                         float2 (&acc)[R]) {
  constexpr int ROW4 = (DP + 2) / 2;
  const float4* rows = reinterpret_cast<const float4*>(chunk);
#pragma unroll 1
  for (int j = 0; j < n; j += 2) {
    float4 a[ROW4], b[ROW4];
    const bool two = j + 1 < n;
    load_row<DP>(a, rows + j * ROW4);
    load_row<DP>(b, rows + (two ? j + 1 : j) * ROW4);
    const float2 wb = two ? hi(b[ROW4 - 1]) : make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float2 ta = row_dot<DP, R>(a, x[r], IS_K ? A[r] : 0.f);
      const float2 tb = row_dot<DP, R>(b, x[r], IS_K ? A[r] : 0.f);
      acc[r] = fma2(IS_K ? ex2_2(ta) : cos_2(ta), hi(a[ROW4 - 1]), acc[r]);
      acc[r] = fma2(IS_K ? ex2_2(tb) : cos_2(tb), wb, acc[r]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward over one chunk: Q_d += t.x row_d.x + t.y row_d.y with
//   features:  t = w' cos(theta + pi/2) = -w' sin(theta)          (A carries the pi/2)
//   inducing:  t = ln2 nu' 2^(theta)  (row weight already holds ln2 nu'), Es += t
// so that  g_k d f_k / d x_d = g_k (Q_d + 2 c_d x_d Es).  Q uses scalar FFMA (same pipe cycles as FFMA2,
// half the registers).
// ---------------------------------------------------------------------------------------------
template <int DP, int R, bool IS_K>
__device__ __forceinline__ void row_bwd_one(const float4 (&v)[(DP + 2) / 2], float2 wgt, const float (&x)[DP], float add, float (&Q)[DP],
                                            float& Es) {
  const float2 th = row_dot<DP, R>(v, x, add);
  const float2 t = mul2(wgt, IS_K ? ex2_2(th) : cos_2(th));
  if (IS_K) Es += t.x + t.y;
#pragma unroll
  for (int i = 0; i < DP / 2; ++i) {
    Q[2 * i] = fmaf(t.x, v[i].x, Q[2 * i]);
    Q[2 * i] = fmaf(t.y, v[i].y, Q[2 * i]);
    Q[2 * i + 1] = fmaf(t.x, v[i].z, Q[2 * i + 1]);
    Q[2 * i + 1] = fmaf(t.y, v[i].w, Q[2 * i + 1]);
  }
}

template <int DP, int R, bool IS_K>
__device__ __forceinline__ void rows_bwd(const float* __restrict__ chunk, int n, const float (&x)[R][DP], const float (&A)[R],
                                         float (&Q)[R][DP], float (&Es)[R]) {
  constexpr int ROW4 = (DP + 2) / 2;
  const float4* rows = reinterpret_cast<const float4*>(chunk);
  if constexpr (DP <= 8) {
#pragma unroll 1
    for (int j = 0; j < n; j += 2) {
      float4 a[ROW4], b[ROW4];
      const bool two = j + 1 < n;
      load_row<DP>(a, rows + j * ROW4);
      load_row<DP>(b, rows + (two ? j + 1 : j) * ROW4);
      const float2 wb = two ? hi(b[ROW4 - 1]) : make_float2(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        row_bwd_one<DP, R, IS_K>(a, hi(a[ROW4 - 1]), x[r], A[r], Q[r], Es[r]);
        row_bwd_one<DP, R, IS_K>(b, wb, x[r], A[r], Q[r], Es[r]);
      }
    }
  } else {  // wide rows: one row per iteration keeps the register footprint under the 3-CTA/SM budget
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
      float4 a[ROW4];
      load_row<DP>(a, rows + j * ROW4);
#pragma unroll
      for (int r = 0; r < R; ++r) row_bwd_one<DP, R, IS_K>(a, hi(a[ROW4 - 1]), x[r], A[r], Q[r], Es[r]);
    }
  }
}

// shared-memory carve-up of the sweep kernels:
//   [mbarriers + producer state (128 B) | kPipeStages x stage | headers D_out x HDR | xs DP x R x threads |
//    (bwd) dx DP x R x threads | dell D_out x DP | dvar D_out]
// xs / dx hold per-thread vectors as [d][r][thread] (conflict free); they let the solver glue run as small
// rolled loops instead of DP x R unrolled register code.
struct SweepSmem {
  uint64_t* bars;
  float* stages;
  float* hdr;
  float* pmax;   // [D_out] max |row coefficient| of this sample's rows (scale of the fp16 dot products)
  float* xs;
  float* dx;
  float* dell;
  float* dvar;
};
template <int DP, int R>
__device__ __forceinline__ SweepSmem carve_smem(float* smem, const RbfGeom& g) {
  SweepSmem s;
  s.bars = reinterpret_cast<uint64_t*>(smem);
  s.stages = smem + 32;
  s.hdr = s.stages + kPipeStages * g.stage_floats;
  s.pmax = s.hdr + g.D_out * g.hdr_floats;
  s.xs = s.pmax + (g.D_out + 3) / 4 * 4;
  s.dx = s.xs + DP * R * blockDim.x;
  s.dell = s.dx + DP * R * blockDim.x;
  s.dvar = s.dell + g.D_out * DP;
  return s;
}

template <int DP, int R>
__device__ __forceinline__ void sweep_setup(SweepSmem& sm, ChunkPipe& pipe, const RbfGeom& g, const float* packed, long total, bool bwd) {
  const int l = blockIdx.y;
  const float* hdr = rbf_hdr_ptr(packed, g, l);
  for (int i = threadIdx.x; i < g.D_out * g.hdr_floats; i += blockDim.x) sm.hdr[i] = hdr[i];
  for (int i = threadIdx.x; i < g.D_out; i += blockDim.x) sm.pmax[i] = rbf_maxabs_ptr(packed, g, l)[i];
  for (int i = threadIdx.x; i < DP * R * blockDim.x; i += blockDim.x) sm.xs[i] = 0.f;  // padded components stay 0
  if (bwd)
    for (int i = threadIdx.x; i < g.D_out * (DP + 1); i += blockDim.x) sm.dell[i] = 0.f;
  // init() contains the __syncthreads that publishes the writes above
  pipe.init(sm.stages, sm.bars, rbf_rows_ptr(packed, g, l), g.cg, total);
}

template <int DP, int R>
__device__ __forceinline__ void load_x(const SweepSmem& sm, float (&x)[R][DP]) {
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int d = 0; d < DP; ++d) x[r][d] = sm.xs[(d * R + r) * blockDim.x + threadIdx.x];
}

// one full forward evaluation of output k: streams the chunks of k, returns prior and update parts
template <int DP, int R>
__device__ __forceinline__ void eval_fwd_k(ChunkPipe& pipe, const RbfGeom& g, long total, const float* hdr_k, const float (&x)[R][DP], float (&fp)[R],
                                           float (&fu)[R]) {
  float A[R];
#pragma unroll
  for (int r = 0; r < R; ++r) A[r] = 0.f;
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    const float c = hdr_k[d];
#pragma unroll
    for (int r = 0; r < R; ++r) A[r] = fmaf(c * x[r][d], x[r][d], A[r]);
  }
  float2 acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
  for (int c = 0; c < g.NCs; ++c) {
    const float* chunk = pipe.acquire(g.cg);
    rows_fwd<DP, R, false>(chunk, min(g.RCs, g.SP2 - c * g.RCs), x, A, acc);
    pipe.release(g.cg, total);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    fp[r] = acc[r].x + acc[r].y;
    acc[r] = make_float2(0.f, 0.f);
  }
  for (int c = 0; c < g.NCm; ++c) {
    const float* chunk = pipe.acquire(g.cg);
    rows_fwd<DP, R, true>(chunk, min(g.RCm, g.MP2 - c * g.RCm), x, A, acc);
    pipe.release(g.cg, total);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) fu[r] = (acc[r].x + acc[r].y) * kInvLn2;  // inducing-row weights carry ln2
}

// one VJP of output k: dxk[r][d] = g_k d f_k / d x_d
template <int DP, int R>
__device__ __forceinline__ void eval_bwd_k(ChunkPipe& pipe, const RbfGeom& g, long total, const float* hdr_k, const float (&x)[R][DP],
                                           const float (&gk)[R], float (&dxk)[R][DP]) {
  float A[R], Ap[R], Es[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    A[r] = 0.f;
    Ap[r] = kHalfPi;
    Es[r] = 0.f;
  }
#pragma unroll
  for (int d = 0; d < DP; ++d) {
    const float c = hdr_k[d];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      A[r] = fmaf(c * x[r][d], x[r][d], A[r]);
      dxk[r][d] = 0.f;
    }
  }
  for (int c = 0; c < g.NCs; ++c) {
    const float* chunk = pipe.acquire(g.cg);
    rows_bwd<DP, R, false>(chunk, min(g.RCs, g.SP2 - c * g.RCs), x, Ap, dxk, Es);
    pipe.release(g.cg, total);
  }
  for (int c = 0; c < g.NCm; ++c) {
    const float* chunk = pipe.acquire(g.cg);
    rows_bwd<DP, R, true>(chunk, min(g.RCm, g.MP2 - c * g.RCm), x, A, dxk, Es);
    pipe.release(g.cg, total);
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int d = 0; d < DP; ++d) dxk[r][d] = gk[r] * fmaf(2.f * hdr_k[d] * x[r][d], Es[r], dxk[r][d]);
}

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


// ---------------------------------------------------------------------------------------------
// policy glue: what the generic sweep kernels (sweep.cuh) call
// ---------------------------------------------------------------------------------------------
// =============================================================================================
// all-k VJP at one state evaluation (stage input already staged in sm.xs): element (k, s) of the
// upstream gradient / f / prior part sits at base[k * kstride + s * sstride].  Leaves sum_k dxk in
// sm.dx and folds the lengthscale / variance statistics.
// =============================================================================================
template <int DP, int R>
__device__ __forceinline__ void vjp_all_k(ChunkPipe& pipe, const RbfGeom& g, long total, const SweepSmem& sm, const States<R>& st,
                                          const float* gvec, const float* fvec, const float* fpvec, long kstride, long sstride) {
  float x[R][DP];
  load_x<DP, R>(sm, x);
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int d = 0; d < DP; ++d) GPODE_XS(sm.dx, d, r) = 0.f;
  for (int k = 0; k < g.D_out; ++k) {
    float gk[R], fk[R], fpk[R], dxk[R][DP];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long at = k * kstride + st.s[r] * sstride;
      gk[r] = st.ok[r] ? gvec[at] : 0.f;
      fk[r] = fvec[at];
      fpk[r] = fpvec[at];
    }
    eval_bwd_k<DP, R>(pipe, g, total, sm.hdr + k * g.hdr_floats, x, gk, dxk);
    rbf_fold_stats<DP, R>(x, dxk, gk, fk, fpk, st.ok, sm.dell, sm.dvar, k);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int d = 0; d < DP; ++d) GPODE_XS(sm.dx, d, r) += dxk[r][d];
  }
}


template <int DP_, int R_>
struct RbfPolicy {
  static constexpr int DP = DP_;
  static constexpr int R = R_;
  static constexpr int kThreads = 256;   // compiled for <= 128 registers: 512 resident threads per SM in any block size
  static constexpr int kMinBlocks = 2;
  static constexpr int kStateThreads = 0;   // every thread of the CTA owns states
  static constexpr int kXsStride = 0;       // staging buffers xs / dx are strided by the block size
  // reverse sweep at D > 8: 128 threads x 3 CTAs (<= 168 registers) -- measured 10% faster than the 128-register build
  static constexpr int kThreadsBwd = DP_ <= 8 ? 256 : 128;
  static constexpr int kMinBlocksBwd = DP_ <= 8 ? 2 : 3;
  using Geom = RbfGeom;
  using Accum = RbfAccum;
  using Smem = SweepSmem;

  __device__ static __forceinline__ Smem carve(float* smem, const Geom& g) { return carve_smem<DP, R>(smem, g); }
  __device__ static __forceinline__ void finish(Smem&) {}
  __device__ static __forceinline__ long setup(Smem& sm, ChunkPipe& pipe, const Geom& g, const float* packed, long n_evals, bool bwd) {
    const long total = n_evals * g.D_out * (g.NCs + g.NCm);
    sweep_setup<DP, R>(sm, pipe, g, packed, total, bwd);
    return total;
  }
  template <class Store>
  __device__ static __forceinline__ void eval_fwd(ChunkPipe& pipe, const Geom& g, long total, const Smem& sm, Store&& store) {
    float x[R][DP];
    load_x<DP, R>(sm, x);
    for (int k = 0; k < g.D_out; ++k) {
      float fp[R], fu[R];
      eval_fwd_k<DP, R>(pipe, g, total, sm.hdr + k * g.hdr_floats, x, fp, fu);
      store(k, fp, fu);
    }
  }
  __device__ static __forceinline__ void vjp(ChunkPipe& pipe, const Geom& g, long total, const Smem& sm, const States<R>& st,
                                             const float* gvec, const float* fvec, const float* fpvec, long kstride, long sstride) {
    vjp_all_k<DP, R>(pipe, g, total, sm, st, gvec, fvec, fpvec, kstride, sstride);
  }
  __device__ static __forceinline__ void flush(const Smem& sm, const Geom& g, const Accum& acc) {
    for (int i = threadIdx.x; i < g.D_out * DP; i += blockDim.x) atomicAdd(&acc.dell_x[i], sm.dell[i]);
    for (int i = threadIdx.x; i < g.D_out; i += blockDim.x) atomicAdd(&acc.dvar[i], sm.dvar[i]);
  }
};

// =============================================================================================
// Forward evaluation on the warp-level tensor path for D > 8 (measurements: rbf_pgrad_mma.cuh).
// A warp owns its 64 states (R = 2) as 4 MMA row tiles of 16.  theta = offset (+ A_k(x)) + x . row is ONE k-step of
// mma.sync m16n8k16 with a two-way fp16 split (head + remainder: 22 bits, three products lo*hi + hi*lo + hi*hi, fp32
// accumulate) -- half the tensor instructions of the 3xTF32 form (6 x m16n8k8), measured 18.8 vs 24.8 ms.  fp16 has no
// exponent headroom, so both operands are scaled by exact powers of two first: every state row by 2^-a (a from its largest
// |x_d|), the rows of output k by 2^-b (b from the largest |coefficient| of that k, found by the pack kernel); the product
// is un-scaled and the offsets added by one FFMA on the C fragment.  Overflow cannot occur, small values degrade like
// block floating point (absolute, not relative, to the row maximum).  The transcendental and the weight run on the C
// fragment; the sum over the rows is a register accumulator per (tile, row half), reduced over the 4 column lanes per k.
// =============================================================================================
__device__ __forceinline__ void mma_tf32_sweep(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16_sweep(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16_sweep(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo_, float hi_) {   // lo_ -> bits 0..15 (the even k index)
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo_, hi_);
  return *reinterpret_cast<const uint32_t*>(&v);
}
// (v0, v1) -> packed fp16 heads and packed fp16 remainders.  The remainders are stored scaled by 2^11 (exact) so that they stay
// fp16-normal wherever the head is; the cross products are accumulated separately and folded in with 2^-11 (kLoScale).
constexpr float kLoUp = 2048.f, kLoScale = 1.f / 2048.f;
__device__ __forceinline__ void split_h2(float v0, float v1, uint32_t& hi_, uint32_t& lo_) {
  const __half2 h = __floats2half2_rn(v0, v1);
  const __half2 l = __floats2half2_rn((v0 - __low2float(h)) * kLoUp, (v1 - __high2float(h)) * kLoUp);
  hi_ = *reinterpret_cast<const uint32_t*>(&h);
  lo_ = *reinterpret_cast<const uint32_t*>(&l);
}
// exact power-of-two scale s (and 1 / s): 1 while 2^-6 <= m < 2^6 (the common case: fp16 heads + remainders then carry
// >= 21 bits of every element above 2^-11 of the maximum), otherwise the scale that brings m into [1/2, 1) -- block floating
// point: no overflow for large operands, no loss of the remainders for small ones
__device__ __forceinline__ void pow2_scales(float m, float& s, float& inv) {
  unsigned be = (__float_as_uint(m) >> 23) & 255u;
  be = be > 252u ? 252u : be;
  const unsigned bs = (be >= 121u && be < 133u) || be == 0u ? 127u : 253u - be;
  s = __uint_as_float(bs << 23);
  inv = __uint_as_float((254u - bs) << 23);
}

// pieces shared by the forward and reverse tensor-path policies
template <int DP>
struct RbfMmaCommon {
  static constexpr int R = 2;
  static constexpr int NT = 4;   // 16-state tiles per warp (32 lanes x R states)

  // A (16 states x 16 dims as fp16 pairs): a[h + 2 j] = s_row x[row gq + 8 h][dims 2 tq + 8 j, 2 tq + 8 j + 1]; state jj of the warp =
  // (r = jj / 32, lane jj % 32).  inv[t][h] = 1 / s_row.
  __device__ static __forceinline__ void build_A(const SweepSmem& sm, int warp, int gq, int tq, uint32_t (&Ah)[NT][4], uint32_t (&Al)[NT][4],
                                                 float (&inv)[NT][2]) {
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int jj = 16 * t + gq + 8 * h;
        const int r = jj >> 5, src = warp * 32 + (jj & 31);
        float v[4], mx = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int d = 2 * tq + 8 * j + i;
            v[2 * j + i] = d < DP ? sm.xs[(d * R + r) * blockDim.x + src] : 0.f;
            mx = fmaxf(mx, fabsf(v[2 * j + i]));
          }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float sa;
        pow2_scales(mx, sa, inv[t][h]);
#pragma unroll
        for (int j = 0; j < 2; ++j) split_h2(v[2 * j] * sa, v[2 * j + 1] * sa, Ah[t][h + 2 * j], Al[t][h + 2 * j]);
      }
  }
  // A_k(x) = sum_d c_kd x_d^2 of every state row: computed thread <-> state from the staged states, handed to the row lanes
  // through a 64-float per-warp scratch
  __device__ static __forceinline__ void quad_A(const SweepSmem& sm, const float* hdr_k, float* scratch, int lane, int gq, float (&Ak)[NT][2]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        const float xv = GPODE_XS(sm.xs, d, r);
        s = fmaf(hdr_k[d] * xv, xv, s);
      }
      scratch[r * 32 + lane] = s;
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int h = 0; h < 2; ++h) Ak[t][h] = scratch[16 * t + gq + 8 * h];
    __syncwarp();
  }
  // B fragments of one block of 8 rows (4 pair rows at `rows`), scaled by sb: column n = gq <-> unit (pair row gq / 2, parity gq & 1);
  // b[j] holds dims 2 tq + 8 j, 2 tq + 8 j + 1
  template <bool CHECK, bool SCALED>
  __device__ static __forceinline__ void build_B(const float* __restrict__ rows, int nvalid, float sb, int gq, int tq, uint32_t (&bh)[2], uint32_t (&bl)[2]) {
    constexpr int ROWF = rbf_row_floats(DP);
    const bool okb = !CHECK || (gq >> 1) < nvalid;
    const float* rowp = rows + (gq >> 1) * ROWF + (gq & 1);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int d = 2 * tq + 8 * j;
      const float v0 = (okb && d < DP) ? rowp[2 * d] : 0.f, v1 = (okb && d + 1 < DP) ? rowp[2 * d + 2] : 0.f;
      if constexpr (SCALED) split_h2(v0 * sb, v1 * sb, bh[j], bl[j]);
      else split_h2(v0, v1, bh[j], bl[j]);
    }
  }
  // theta of one tile: init + u_row (x' . row').  SCALED = false (every scale of this warp and output is 1, the common case): the
  // offsets are simply the initial accumulator
  template <bool SCALED>
  __device__ static __forceinline__ void theta_tile(const uint32_t (&Ah)[4], const uint32_t (&Al)[4], const uint32_t (&bh)[2], const uint32_t (&bl)[2],
                                                    float u0, float u1, const float (&init)[4], float (&th)[4]) {
    float x2[4] = {0.f, 0.f, 0.f, 0.f};   // cross terms lo' * hi + hi * lo' (x 2^11)
    mma_f16_sweep(x2, Al, bh[0], bh[1]);
    mma_f16_sweep(x2, Ah, bl[0], bl[1]);
    if constexpr (SCALED) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      mma_f16_sweep(c, Ah, bh[0], bh[1]);
      th[0] = fmaf(x2[0], u0 * kLoScale, fmaf(c[0], u0, init[0]));
      th[1] = fmaf(x2[1], u0 * kLoScale, fmaf(c[1], u0, init[1]));
      th[2] = fmaf(x2[2], u1 * kLoScale, fmaf(c[2], u1, init[2]));
      th[3] = fmaf(x2[3], u1 * kLoScale, fmaf(c[3], u1, init[3]));
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) th[i] = init[i];
      mma_f16_sweep(th, Ah, bh[0], bh[1]);
#pragma unroll
      for (int i = 0; i < 4; ++i) th[i] = fmaf(x2[i], kLoScale, th[i]);
    }
  }
  // true when every un-scaling factor of this warp is exactly 1 (warp uniform)
  __device__ static __forceinline__ bool all_unit(const float (&u)[NT][2]) {
    bool one = true;
#pragma unroll
    for (int t = 0; t < NT; ++t) one = one && u[t][0] == 1.f && u[t][1] == 1.f;
    return __all_sync(0xffffffffu, one);
  }
};

template <int DP_>
struct RbfMmaFwdPolicy : RbfPolicy<DP_, 2> {
  static constexpr int DP = DP_;
  static constexpr int R = 2;
  static constexpr int NT = 4;
  using C = RbfMmaCommon<DP_>;

  // one block of 8 features / inducing points against the warp's 4 state tiles; CHECK: the block may run past the n valid pair
  // rows of the chunk (tail block only)
  template <bool IS_K, bool CHECK, bool SCALED>
  __device__ static __forceinline__ void block_mma(const float* __restrict__ rows, int nvalid, float sb, const uint32_t (&Ah)[NT][4],
                                                   const uint32_t (&Al)[NT][4], const float (&u)[NT][2], const float (&Ak)[NT][2], float (&acc)[NT][2],
                                                   int gq, int tq) {
    constexpr int ROWF = rbf_row_floats(DP);
    uint32_t bh[2], bl[2];
    C::template build_B<CHECK, SCALED>(rows, nvalid, sb, gq, tq, bh, bl);
    // offsets and weights of the C columns 2 tq, 2 tq + 1 = both parities of pair row tq
    float2 off = make_float2(0.f, 0.f), wgt = make_float2(0.f, 0.f);
    if (!CHECK || tq < nvalid) {
      off = *reinterpret_cast<const float2*>(rows + tq * ROWF + 2 * DP);
      wgt = *reinterpret_cast<const float2*>(rows + tq * ROWF + 2 * DP + 2);
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      float init[4], th[4];
      init[0] = IS_K ? Ak[t][0] + off.x : off.x;
      init[1] = IS_K ? Ak[t][0] + off.y : off.y;
      init[2] = IS_K ? Ak[t][1] + off.x : off.x;
      init[3] = IS_K ? Ak[t][1] + off.y : off.y;
      C::template theta_tile<SCALED>(Ah[t], Al[t], bh, bl, u[t][0], u[t][1], init, th);
      const float v0 = IS_K ? ex2_approx(th[0]) : __cosf(th[0]), v1 = IS_K ? ex2_approx(th[1]) : __cosf(th[1]);
      const float v2 = IS_K ? ex2_approx(th[2]) : __cosf(th[2]), v3 = IS_K ? ex2_approx(th[3]) : __cosf(th[3]);
      acc[t][0] = fmaf(v0, wgt.x, acc[t][0]);
      acc[t][0] = fmaf(v1, wgt.y, acc[t][0]);
      acc[t][1] = fmaf(v2, wgt.x, acc[t][1]);
      acc[t][1] = fmaf(v3, wgt.y, acc[t][1]);
    }
  }
  template <bool IS_K>
  __device__ static __forceinline__ void rows_mma(const float* __restrict__ chunk, int n, bool plain, float sb, const uint32_t (&Ah)[NT][4],
                                                  const uint32_t (&Al)[NT][4], const float (&u)[NT][2], const float (&Ak)[NT][2], float (&acc)[NT][2], int gq,
                                                  int tq) {
    constexpr int ROWF = rbf_row_floats(DP);
    const int nfull = n >> 2;
    if (plain) {
#pragma unroll 1
      for (int blk = 0; blk < nfull; ++blk) block_mma<IS_K, false, false>(chunk + blk * 4 * ROWF, 4, sb, Ah, Al, u, Ak, acc, gq, tq);
      if (n & 3) block_mma<IS_K, true, false>(chunk + nfull * 4 * ROWF, n & 3, sb, Ah, Al, u, Ak, acc, gq, tq);
    } else {
#pragma unroll 1
      for (int blk = 0; blk < nfull; ++blk) block_mma<IS_K, false, true>(chunk + blk * 4 * ROWF, 4, sb, Ah, Al, u, Ak, acc, gq, tq);
      if (n & 3) block_mma<IS_K, true, true>(chunk + nfull * 4 * ROWF, n & 3, sb, Ah, Al, u, Ak, acc, gq, tq);
    }
  }

  template <class Store>
  __device__ static __forceinline__ void eval_fwd(ChunkPipe& pipe, const RbfGeom& g, long total, const SweepSmem& sm, Store&& store) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gq = lane >> 2, tq = lane & 3;
    uint32_t Ah[NT][4], Al[NT][4];
    float inv[NT][2];
    C::build_A(sm, warp, gq, tq, Ah, Al, inv);
    float* res = sm.dx + warp * 128;   // per-warp scratch [2][64]: prior part / update part of the warp's states
    for (int k = 0; k < g.D_out; ++k) {
      float sb, isb, Ak[NT][2], u[NT][2], acc[NT][2];
      pow2_scales(sm.pmax[k], sb, isb);
      C::quad_A(sm, sm.hdr + k * g.hdr_floats, res, lane, gq, Ak);
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          u[t][h] = inv[t][h] * isb;
          acc[t][h] = 0.f;
        }
      const bool plain = C::all_unit(u);
      for (int c = 0; c < g.NCs; ++c) {
        const float* chunk = pipe.acquire(g.cg);
        rows_mma<false>(chunk, min(g.RCs, g.SP2 - c * g.RCs), plain, sb, Ah, Al, u, Ak, acc, gq, tq);
        pipe.release(g.cg, total);
      }
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float s = acc[t][h];
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          if (tq == 0) res[16 * t + gq + 8 * h] = s;
          acc[t][h] = 0.f;
        }
      for (int c = 0; c < g.NCm; ++c) {
        const float* chunk = pipe.acquire(g.cg);
        rows_mma<true>(chunk, min(g.RCm, g.MP2 - c * g.RCm), plain, sb, Ah, Al, u, Ak, acc, gq, tq);
        pipe.release(g.cg, total);
      }
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float s = acc[t][h];
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          if (tq == 0) res[64 + 16 * t + gq + 8 * h] = s * kInvLn2;   // inducing-row weights carry ln2
        }
      __syncwarp();
      float fp[R], fu[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        fp[r] = res[r * 32 + lane];
        fu[r] = res[64 + r * 32 + lane];
      }
      __syncwarp();
      store(k, fp, fu);
    }
  }
};

// =============================================================================================
// Reverse-sweep VJP on the warp-level tensor path for D > 8: same tiling as RbfMmaFwdPolicy.  Per block of 8 rows and state
// tile: theta (3 fp16 MMAs, as in the forward), t = weight * (-sin | 2^theta) on the C fragment, and the second product
// Q[state, d] += t P[row, d] with the C fragment reused as the A fragment (row order permuted consistently in B, no shuffles):
// t spans the dynamic range of the exponentials, so it keeps the fp32 exponent -- head x head is one TF32 m16n8k8, the two
// cross terms rem(t) P + t rem(P) share ONE bf16 m16n8k16 (k = 8 rows x {remainder, value}; bf16 has the fp32 exponent and
// the cross terms only need 8 of their bits: 2^-19 relative overall) -- 4 MMAs per tile instead of the 6 of 3xTF32.  These
// are the two D-length contractions of the FFMA kernel.  dx_k = g_k (Q + 2 c_k x Es) is folded into the shared dx / lengthscale
// statistics once per output dimension.
// =============================================================================================
template <int DP_>
struct RbfMmaBwdPolicy : RbfPolicy<DP_, 2> {
  static constexpr int DP = DP_;
  static constexpr int R = 2;
  static constexpr int NT = 4;
  static constexpr int NB = 2;   // 8-column blocks of the input dimension in the second product
  using C = RbfMmaCommon<DP_>;

  template <bool IS_K, bool CHECK, bool SCALED>
  __device__ static __forceinline__ void block_bwd(const float* __restrict__ rows, int nvalid, float sb, const uint32_t (&Ah)[NT][4],
                                                   const uint32_t (&Al)[NT][4], const float (&u)[NT][2], const float (&Ak)[NT][2], float (&Q)[NT][NB][4],
                                                   float (&Es)[NT][2], int gq, int tq) {
    constexpr int ROWF = rbf_row_floats(DP);
    const bool okc = !CHECK || tq < nvalid;
    uint32_t bh[2], bl[2];
    C::template build_B<CHECK, SCALED>(rows, nvalid, sb, gq, tq, bh, bl);
    float2 off = make_float2(0.f, 0.f), wgt = make_float2(0.f, 0.f);
    if (okc) {
      off = *reinterpret_cast<const float2*>(rows + tq * ROWF + 2 * DP);
      wgt = *reinterpret_cast<const float2*>(rows + tq * ROWF + 2 * DP + 2);
    }
    // B of the second product (3xTF32: t has the dynamic range of the exponentials): MMA k index tq <-> row unit 2 tq, k = tq + 4 <-> unit
    // 2 tq + 1 (both parities of pair row tq); column gq <-> dim gq + 8 nb
    uint32_t ph[NB][2], pc[NB][2];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int d = gq + 8 * nb;
      float2 pv = make_float2(0.f, 0.f);
      if (okc && d < DP) pv = *reinterpret_cast<const float2*>(rows + tq * ROWF + 2 * d);
      ph[nb][0] = __float_as_uint(pv.x) & 0xFFFFE000u;
      ph[nb][1] = __float_as_uint(pv.y) & 0xFFFFE000u;
      // correction operand (bf16 k16): k 2 tq, 2 tq + 1 <-> P of units 2 tq, 2 tq + 1; k + 8 <-> their TF32 remainders
      pc[nb][0] = pack_bf16(pv.x, pv.y);
      pc[nb][1] = pack_bf16(pv.x - __uint_as_float(ph[nb][0]), pv.y - __uint_as_float(ph[nb][1]));
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      float init[4], th[4];
      init[0] = IS_K ? Ak[t][0] + off.x : off.x + kHalfPi;    // cos(theta + pi/2) = -sin(theta)
      init[1] = IS_K ? Ak[t][0] + off.y : off.y + kHalfPi;
      init[2] = IS_K ? Ak[t][1] + off.x : off.x + kHalfPi;
      init[3] = IS_K ? Ak[t][1] + off.y : off.y + kHalfPi;
      C::template theta_tile<SCALED>(Ah[t], Al[t], bh, bl, u[t][0], u[t][1], init, th);
      // C fragment: [0] (row gq, unit 2 tq), [1] (gq, 2 tq + 1), [2] (gq + 8, 2 tq), [3] (gq + 8, 2 tq + 1)
      const float t0 = wgt.x * (IS_K ? ex2_approx(th[0]) : __cosf(th[0])), t1 = wgt.y * (IS_K ? ex2_approx(th[1]) : __cosf(th[1]));
      const float t2 = wgt.x * (IS_K ? ex2_approx(th[2]) : __cosf(th[2])), t3 = wgt.y * (IS_K ? ex2_approx(th[3]) : __cosf(th[3]));
      if (IS_K) {
        Es[t][0] += t0 + t1;
        Es[t][1] += t2 + t3;
      }
      // A fragment of the second product: a0 (row gq, k tq) = t0, a1 (row gq + 8, k tq) = t2, a2 (gq, tq + 4) = t1, a3 = t3
      const float v4[4] = {t0, t2, t1, t3};
      uint32_t ah[4], ac[4];
      float al[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ah[i] = __float_as_uint(v4[i]) & 0xFFFFE000u;
        al[i] = v4[i] - __uint_as_float(ah[i]);
      }
      // correction fragment: (row gq | gq + 8, k 2 tq, 2 tq + 1) = remainders of t, (k + 8) = t itself
      ac[0] = pack_bf16(al[0], al[2]);
      ac[1] = pack_bf16(al[1], al[3]);
      ac[2] = pack_bf16(t0, t1);
      ac[3] = pack_bf16(t2, t3);
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        mma_bf16_sweep(Q[t][nb], ac, pc[nb][0], pc[nb][1]);
        mma_tf32_sweep(Q[t][nb], ah, ph[nb][0], ph[nb][1]);
      }
    }
  }
  template <bool IS_K>
  __device__ static __forceinline__ void rows_bwd_mma(const float* __restrict__ chunk, int n, bool plain, float sb, const uint32_t (&Ah)[NT][4],
                                                      const uint32_t (&Al)[NT][4], const float (&u)[NT][2], const float (&Ak)[NT][2], float (&Q)[NT][NB][4],
                                                      float (&Es)[NT][2], int gq, int tq) {
    constexpr int ROWF = rbf_row_floats(DP);
    const int nfull = n >> 2;
    if (plain) {
#pragma unroll 1
      for (int blk = 0; blk < nfull; ++blk) block_bwd<IS_K, false, false>(chunk + blk * 4 * ROWF, 4, sb, Ah, Al, u, Ak, Q, Es, gq, tq);
      if (n & 3) block_bwd<IS_K, true, false>(chunk + nfull * 4 * ROWF, n & 3, sb, Ah, Al, u, Ak, Q, Es, gq, tq);
    } else {
#pragma unroll 1
      for (int blk = 0; blk < nfull; ++blk) block_bwd<IS_K, false, true>(chunk + blk * 4 * ROWF, 4, sb, Ah, Al, u, Ak, Q, Es, gq, tq);
      if (n & 3) block_bwd<IS_K, true, true>(chunk + nfull * 4 * ROWF, n & 3, sb, Ah, Al, u, Ak, Q, Es, gq, tq);
    }
  }

  __device__ static __forceinline__ void vjp(ChunkPipe& pipe, const RbfGeom& g, long total, const SweepSmem& sm, const States<R>& st,
                                             const float* gvec, const float* fvec, const float* fpvec, long kstride, long sstride) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gq = lane >> 2, tq = lane & 3;
    uint32_t Ah[NT][4], Al[NT][4];
    float inv[NT][2];
    C::build_A(sm, warp, gq, tq, Ah, Al, inv);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int d = 0; d < DP; ++d) GPODE_XS(sm.dx, d, r) = 0.f;
    float* gs = sm.dvar + g.D_out + warp * 128;   // per-warp scratch: [64] upstream gradient of the warp's states, [64] A_k hand-over
    for (int k = 0; k < g.D_out; ++k) {
      const float* hdr_k = sm.hdr + k * g.hdr_floats;
      // thread <-> state part: upstream gradient, variance statistic
      {
        float v = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const long at = k * kstride + st.s[r] * sstride;
          const float gk = st.ok[r] ? gvec[at] : 0.f;
          v += gk * (fvec[at] - 0.5f * fpvec[at]);
          gs[r * 32 + lane] = gk;
        }
        v = warp_sum(v);
        if (lane == 0) atomicAdd(&sm.dvar[k], v);
      }
      float sb, isb, Ak[NT][2], u[NT][2], Es[NT][2], Q[NT][NB][4];
      pow2_scales(sm.pmax[k], sb, isb);
      C::quad_A(sm, hdr_k, gs + 64, lane, gq, Ak);     // (contains the __syncwarp that publishes gs)
#pragma unroll
      for (int t = 0; t < NT; ++t) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          u[t][h] = inv[t][h] * isb;
          Es[t][h] = 0.f;
        }
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
          for (int i = 0; i < 4; ++i) Q[t][nb][i] = 0.f;
      }
      const bool plain = C::all_unit(u);
      for (int c = 0; c < g.NCs; ++c) {
        const float* chunk = pipe.acquire(g.cg);
        rows_bwd_mma<false>(chunk, min(g.RCs, g.SP2 - c * g.RCs), plain, sb, Ah, Al, u, Ak, Q, Es, gq, tq);
        pipe.release(g.cg, total);
      }
      for (int c = 0; c < g.NCm; ++c) {
        const float* chunk = pipe.acquire(g.cg);
        rows_bwd_mma<true>(chunk, min(g.RCm, g.MP2 - c * g.RCm), plain, sb, Ah, Al, u, Ak, Q, Es, gq, tq);
        pipe.release(g.cg, total);
      }
      // dx_k[state][d] = g_k (Q + 2 c_d x_d Es) in the C layout (rows gq + 8 h, dims 8 nb + 2 tq + j); fold into dx and the lengthscale statistic
      float stat[NB][2];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) stat[nb][0] = stat[nb][1] = 0.f;
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float es = Es[t][h];
          es += __shfl_xor_sync(0xffffffffu, es, 1);
          es += __shfl_xor_sync(0xffffffffu, es, 2);
          const int jj = 16 * t + gq + 8 * h;
          const int r = jj >> 5, src = warp * 32 + (jj & 31);
          const float gk = gs[jj];
#pragma unroll
          for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int d = 8 * nb + 2 * tq + j;
              if (d < DP) {
                const float xv = sm.xs[(d * R + r) * blockDim.x + src];
                const float dxk = gk * fmaf(2.f * hdr_k[d] * xv, es, Q[t][nb][2 * h + j]);
                sm.dx[(d * R + r) * blockDim.x + src] += dxk;
                stat[nb][j] = fmaf(xv, dxk, stat[nb][j]);
              }
            }
        }
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float v = stat[nb][j];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          const int d = 8 * nb + 2 * tq + j;
          if (gq == 0 && d < DP) atomicAdd(&sm.dell[k * DP + d], v);
        }
      __syncwarp();
    }
  }
};

// =============================================================================================
// parameter gradients: threads <-> inducing-point pairs of one (sample, output dim); the CTA walks a
// chunk of state evaluations staged through shared memory and accumulates in registers:
//   dnu'_m = sum_n g_n E_nm ,  pg_md = sum_n g_n E_nm x_nd
// =============================================================================================
constexpr int kPgBatch = 128;


// PP inducing-point pairs per thread: one broadcast LDS.128 of a staged state feeds 2 PP inducing points (PP = 1 is LDS-bound)
template <int DP, int PP>
__global__ void __launch_bounds__(kPgThreads, (DP <= 8 ? 6 : 3)) k_rbf_pgrad(const RbfPgradArgs a) {
  const RbfGeom& g = a.g;
  constexpr int ROW4 = (DP + 2) / 2;
  constexpr int SROW = ((DP + 2 + 3) / 4) * 4;  // staged state: x[DP], g, A (+pad), 16-byte rows
  constexpr int NV = SROW / 4;
  __shared__ __align__(16) float stage[kPgBatch * SROW];
  __shared__ float s_c[DP];
  const int k = blockIdx.y, l = blockIdx.z;
  const float* hdr = rbf_hdr_ptr(a.packed, g, l) + k * g.hdr_floats;
  if (threadIdx.x < DP) s_c[threadIdx.x] = hdr[threadIdx.x];
  const int per_cta = blockDim.x * PP;
  const int n_mblk = (g.MP2 + per_cta - 1) / per_cta;
  const int chunk_id = blockIdx.x / n_mblk;
  const int j0 = (blockIdx.x - chunk_id * n_mblk) * per_cta + threadIdx.x;  // first inducing pair of this thread
  float2 G[PP][DP], H[PP], dnu[PP], pg[PP][DP];
  bool any = false;
#pragma unroll
  for (int p = 0; p < PP; ++p) {
    const int j = j0 + p * blockDim.x;
    H[p] = make_float2(0.f, 0.f);
    dnu[p] = make_float2(0.f, 0.f);
#pragma unroll
    for (int d = 0; d < DP; ++d) {
      G[p][d] = make_float2(0.f, 0.f);
      pg[p][d] = make_float2(0.f, 0.f);
    }
    if (j < g.MP2) {
      any = true;
      const float4* row = reinterpret_cast<const float4*>(rbf_rows_ptr(a.packed, g, l) +
                                                          (static_cast<size_t>(k) * (g.SP2 + g.MP2) + g.SP2 + j) * g.row_floats);
#pragma unroll
      for (int i = 0; i < DP / 2; ++i) {
        const float4 v = row[i];
        G[p][2 * i] = lo(v);
        G[p][2 * i + 1] = hi(v);
      }
      H[p] = lo(row[ROW4 - 1]);
    }
  }

  const long total = a.n_te * g.N;  // state evaluations of this sample
  const long per = (total + a.chunks - 1) / a.chunks;
  const long e_lo = static_cast<long>(chunk_id) * per;
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
    if (any) {
      const int nb = (e_hi - e0) < kPgBatch ? static_cast<int>(e_hi - e0) : kPgBatch;
#pragma unroll(PP == 1 ? 2 : 1)
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
        float2 ex[PP];
        if constexpr (DP >= 8) {   // two partial sums: halves the dependent FFMA2 chain
          float2 ey[PP];
#pragma unroll
          for (int p = 0; p < PP; ++p) {
            ex[p] = fma2(bc(xv[0]), G[p][0], add2(H[p], bc(xv[DP + 1])));
            ey[p] = mul2(bc(xv[1]), G[p][1]);
          }
#pragma unroll
          for (int d = 2; d < DP; d += 2)
#pragma unroll
            for (int p = 0; p < PP; ++p) {
              ex[p] = fma2(bc(xv[d]), G[p][d], ex[p]);
              ey[p] = fma2(bc(xv[d + 1]), G[p][d + 1], ey[p]);
            }
#pragma unroll
          for (int p = 0; p < PP; ++p) ex[p] = add2(ex[p], ey[p]);
        } else {
#pragma unroll
          for (int p = 0; p < PP; ++p) ex[p] = add2(H[p], bc(xv[DP + 1]));
#pragma unroll
          for (int d = 0; d < DP; ++d)
#pragma unroll
            for (int p = 0; p < PP; ++p) ex[p] = fma2(bc(xv[d]), G[p][d], ex[p]);
        }
#pragma unroll
        for (int p = 0; p < PP; ++p) {
          const float2 ge = mul2(bc(xv[DP]), ex2_2(ex[p]));
          dnu[p] = add2(dnu[p], ge);
#pragma unroll
          for (int d = 0; d < DP; ++d) pg[p][d] = fma2(bc(xv[d]), ge, pg[p][d]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < PP; ++p) {
    const int j = j0 + p * blockDim.x;
    if (j < g.MP2) {
      const size_t base = (static_cast<size_t>(l) * g.D_out + k) * (2 * g.MP2) + 2 * j;
      atomicAdd(&a.acc.dnu[base], dnu[p].x);
      atomicAdd(&a.acc.dnu[base + 1], dnu[p].y);
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        atomicAdd(&a.acc.pg[base * DP + d], pg[p][d].x);
        atomicAdd(&a.acc.pg[(base + 1) * DP + d], pg[p][d].y);
      }
    }
  }
}

// launch-shape heuristic of the sweep kernels.  All sweep kernels are compiled for <= 128 registers (512 resident
// threads per SM), so any block size up to 256 threads can be launched; states per CTA = threads * R.  The model
// below estimates the time of a launch as (CTAs the busiest SM runs back to back) x (work per CTA) x (cost per state of
// the register-blocking factor R: a broadcast LDS.128 feeds R states, small R is LDS-bound) and picks the cheapest shape;
// for a chip-filling problem this amounts to choosing the shape whose last wave is full.
inline void rbf_pick_shape(const RbfGeom& g, bool bwd, int& threads, int& R, int force_R = 0) {
  const int rc[3] = {4, 2, 1};
  double best = 1e300;
  threads = 32;
  R = 1;
  for (int ri = 0; ri < 3; ++ri) {
    const int r = rc[ri];
    if (force_R && r != force_R) continue;
    if (bwd && g.DP > 8 && r > 2) continue;   // not compiled: the reverse sweep at D > 8 holds 2 states per thread at most
    // relative cost per state: forward needs R*2 FMA per parameter float to hide the LDS, the reverse sweep uses each float twice
    const double cost = bwd ? (r == 1 ? 1.25 : 1.0) : (r == 1 ? (g.DP > 8 ? 2.0 : 1.6) : (r == 2 ? (g.DP > 8 ? 1.2 : 1.05) : 1.0));
    for (int t = 32; t <= 256; t += 32) {   // ties go to the smaller block
      const int smem = rbf_smem_bytes(g, t, r, bwd) + 1024;
      if (bwd && g.DP > 8 && t > 128) continue;
      int bps = ((bwd && g.DP > 8) ? 384 : 512) / t;
      const int bps_s = (228 * 1024) / smem;
      if (bps_s < bps) bps = bps_s;
      if (bps > 32) bps = 32;
      if (bps < 1) continue;
      const long per = static_cast<long>(t) * r;
      const long ctas = ((static_cast<long>(g.N) + per - 1) / per) * g.L;
      const long slots = 148L * bps;
      const long waves = (ctas + slots - 1) / slots;
      const long rem = ctas - (waves - 1) * slots;
      const long load = (waves - 1) * bps + (rem + 147) / 148;            // CTAs the busiest SM runs
      const long resident = load < bps ? load : bps;
      const double warps = static_cast<double>(resident) * t / 32.0;
      const double latency = warps >= 12.0 ? 1.0 : (warps >= 8.0 ? 1.08 : 8.0 * 1.08 / warps);   // too few warps: latency bound
      const double time = static_cast<double>(load) * per * cost * latency;
      if (time < best * 0.999) {
        best = time;
        threads = t;
        R = r;
      }
    }
  }
}

}  // namespace gpode
