// Divergence-free (DF) sparse-GP vector field: policy for the generic sweep kernels (sweep.cuh) and the
// parameter-gradient kernels, sm_100a.  Math and packed layout: df.h.
//
// Mapping: like the RBF kernels, threads <-> R states, parameter rows stream through shared memory
// (ChunkPipe, 1-D TMA); every row carries a PAIR of features / inducing points so the inner loops are
// FFMA2 with the state as the scalar-broadcast operand.  The D x D block of exponentials per
// (state, inducing point) makes the update part MUFU.EX2-bound (D^2 MUFU against ~5 D^2 FMA-pipe slots),
// the per-(i,j) constants {k_ij, lc_ij} are broadcast LDS.128 from a shared-memory header.
#pragma once

#include "common.cuh"
#include "df.h"
#include "sweep.cuh"

namespace gpode {

constexpr float kDfHalfPi = 1.5707963267948966f;
constexpr float kTwoLog2e = 2.8853900817779268f;

template <int N4>
__device__ __forceinline__ float2 f2at(const float4 (&v)[N4], int q) {
  return (q & 1) ? hi(v[q >> 1]) : lo(v[q >> 1]);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return fma2(b, bc(-1.f), a); }
// broadcast load of a per-(i,j) constant; volatile so that the D*D loop-invariant loads are NOT hoisted out of the
// inducing-point loop into 4 D^2 registers (they are cheaper as LDS.128 than as spills)
__device__ __forceinline__ float4 lds_const4(const float4* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

struct DfSmem {
  uint64_t* bars;
  float* stages;
  const float4* kc;   // [D*D] {k,k,lc,lc}
  const float* h;     // [D]
  float* xs;
  float* dx;
  float* dell;        // [D*D] then dvar [D], then dc [D*D]
  float* dvar;
  float* dcs;
  float* pacc;        // [MP2][4D]: per inducing pair {dnu even, dnu odd, dZ even, dZ odd}, summed over this CTA's states
  float* xch;         // cluster launches: [2][C][2 D][threads] all-gathered partial sums of one evaluation (double-buffered over evaluations)
  float* loc;         // W > 0: [W][2 D][32] partial sums of the CTA's warps
  int ev;             // running evaluation counter (buffer parity)
};

// Small batches (the reference's own training shapes, e.g. 256 trajectories x 4 samples): a few dozen warps each walking every parameter
// row serially leave the chip idle (config 2: 32 one-warp CTAs).  The ROWS of every streamed chunk are then split over a thread-block
// cluster along grid z (df_cluster(), <= 8 CTAs): every CTA streams all chunks but evaluates only its slice of the rows, the partial
// sums of the field (forward) / of J^T g (reverse sweep) are all-gathered through distributed shared memory with one cluster barrier per
// evaluation, and every CTA runs the tiny solver glue redundantly on identical values -- the generic sweep kernels need no cross-CTA logic.
// Statistics that are sums over rows stay per CTA (each adds its share to the global accumulators); the variance statistic (no row sum)
// is taken by rank 0 only.  With W > 0 (DfPolicy<D, 1, W>, the small-batch instantiation) a CTA owns 32 states and brings W warps that
// split the row slice once more (lane <-> state in every warp; the warps' partial sums meet in shared memory before the all-gather):
// config 2 runs 128 CTAs x 8 warps instead of 32 CTAs x 1 warp.
__device__ __forceinline__ void df_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void df_st_cluster(uint32_t local_saddr, int rank, float v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
// rows [lo, lo + cnt) of a chunk of n rows that cluster rank `rank` of C evaluates
__device__ __forceinline__ void df_slice(int n, int C, int rank, int& lo, int& cnt) {
  lo = n * rank / C;
  cnt = n * (rank + 1) / C - lo;
}

// ---------------------------------------------------------------------------------------------
// prior part, one chunk of n feature rows of block a:  row = {Om'_d}_d<D, b', {B'_c}_c<D, pad  (float2 pairs)
// ---------------------------------------------------------------------------------------------
template <int D, int R>
__device__ __forceinline__ void df_rows_prior_fwd(const float* __restrict__ chunk, int n, const float (&x)[R][D], float2 (&fp)[R][D]) {
  constexpr int ROW4 = D + 1;
  const float4* rows = reinterpret_cast<const float4*>(chunk);
#pragma unroll 1
  for (int j = 0; j < n; ++j) {
    float4 v[ROW4];
#pragma unroll
    for (int i = 0; i < ROW4; ++i) v[i] = rows[j * ROW4 + i];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 th = f2at(v, D);
#pragma unroll
      for (int d = 0; d < D; ++d) th = fma2(bc(x[r][d]), f2at(v, d), th);
      const float2 cs = cos_2(th);
#pragma unroll
      for (int c = 0; c < D; ++c) fp[r][c] = fma2(cs, f2at(v, D + 1 + c), fp[r][c]);
    }
  }
}

// VJP of the prior part: t = -sin(theta') sum_c g_c B'_c ; Q_d += t Om'_d  (scalar FFMA: half the registers)
template <int D, int R>
__device__ __forceinline__ void df_rows_prior_bwd(const float* __restrict__ chunk, int n, const float (&x)[R][D], const float (&g)[R][D],
                                                  float (&Q)[R][D]) {
  constexpr int ROW4 = D + 1;
  const float4* rows = reinterpret_cast<const float4*>(chunk);
#pragma unroll 1
  for (int j = 0; j < n; ++j) {
    float4 v[ROW4];
#pragma unroll
    for (int i = 0; i < ROW4; ++i) v[i] = rows[j * ROW4 + i];
    const float2 b = add2(f2at(v, D), bc(kDfHalfPi));   // cos(th + pi/2) = -sin(th)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 th = b, gc = make_float2(0.f, 0.f);
#pragma unroll
      for (int d = 0; d < D; ++d) th = fma2(bc(x[r][d]), f2at(v, d), th);
#pragma unroll
      for (int c = 0; c < D; ++c) gc = fma2(bc(g[r][c]), f2at(v, D + 1 + c), gc);
      const float2 t = mul2(cos_2(th), gc);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float2 om = f2at(v, d);
        Q[r][d] = fmaf(t.x, om.x, Q[r][d]);
        Q[r][d] = fmaf(t.y, om.y, Q[r][d]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// update part, one chunk of n inducing rows: row = {z_d}_d<D, {nu_i}_i<D  (float2 = two inducing points)
// ---------------------------------------------------------------------------------------------
template <int D, int R>
__device__ __forceinline__ void df_rows_k_fwd(const float* __restrict__ chunk, int n, const DfSmem& sm, const float (&x)[R][D],
                                              float2 (&F)[R][D]) {
  const float4* rows = reinterpret_cast<const float4*>(chunk);
#pragma unroll 1
  for (int m = 0; m < n; ++m) {
    float4 v[D];
#pragma unroll
    for (int i = 0; i < D; ++i) v[i] = rows[m * D + i];
    float2 d[R][D], p[R][D], r2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      r2[r] = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < D; ++k) {
        d[r][k] = sub2(bc(x[r][k]), f2at(v, k));
        r2[r] = fma2(d[r][k], d[r][k], r2[r]);
        p[r][k] = mul2(f2at(v, D + k), d[r][k]);
      }
    }
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float2 s[R], ejj[R];
#pragma unroll
      for (int r = 0; r < R; ++r) s[r] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float4 kc = lds_const4(sm.kc + i * D + j);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float2 e = ex2_2(fma2(r2[r], lo(kc), hi(kc)));
          s[r] = fma2(p[r][i], e, s[r]);
          if (i == j) ejj[r] = e;
        }
      }
      const float hj = sm.h[j];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        F[r][j] = fma2(d[r][j], s[r], F[r][j]);
        F[r][j] = fma2(mul2(f2at(v, D + j), ejj[r]), sub2(bc(hj), r2[r]), F[r][j]);
      }
    }
  }
}

// Cross-lane "transpose reduction": every lane holds K values; afterwards lane q holds the sum over the 32 lanes of value
// xred_index(q) (K <= 32).  Recursive halving: K + log-many SHFLs instead of 5 K for K independent warp sums.
template <int K, int OFF>
struct XRed {
  static __device__ __forceinline__ float run(const float (&v)[K], int lane) {
    constexpr int H = (K + 1) / 2;
    const bool up = (lane & OFF) != 0;
    float w[H];
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const float a = v[i];
      const float b = (i + H < K) ? v[i + H] : 0.f;
      const float send = up ? a : b, keep = up ? b : a;
      w[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
    if constexpr (OFF == 1) {
      return w[0];
    } else {
      return XRed<H, OFF / 2>::run(w, lane);
    }
  }
};
// which of the K values lane `lane` ends up holding (-1: none)
__device__ __forceinline__ int xred_index(int K, int lane) {
  int idx = 0, size = K, valid = K;   // size: (zero-padded) array length at this level, the same in every lane
  for (int off = 16; off > 0; off >>= 1) {
    const int h = (size + 1) / 2;
    if (lane & off) {
      idx += h;
      valid -= h;
    } else {
      valid = valid < h ? valid : h;
    }
    size = h;
  }
  return valid >= 1 ? idx : -1;
}

// VJP of the update part AND its parameter gradients (SURVEY.md Appendix A.6 rearranged; checked in tests/test_df_algebra.py):
//   u_j = g_j d_j, p_i = nu_i d_i, s_j = sum_i p_i e_ij, t_i = sum_j u_j e_ij, W = sum_ij u_j p_i c_ij e_ij,
//   V = sum_j g_j nu_j e_jj (c_jj (h_j - r2) + 2)   ->   dL/dd_k = g_k s_k + nu_k t_k - d_k (W + V)
//   dL/dx_k += dL/dd_k ;  dL/dZ_mk = -sum_n dL/dd_k ;  dL/dnu_mi = sum_n [d_i t_i + g_i e_ii (h_i - r2)]
//   dc'_ij  = sum_{n,m} u_j p_i e_ij (r2 k_ij + 2 log2e) (+ diagonal term)      (lengthscale gradient, finalize kernel)
// with c_ij e_ij = -2 ln2 k_ij e_ij (k_ij is already in a register pair).  The per-inducing-point sums over the states
// of the warp are transpose-reduced with shuffles (4D values per inducing pair) and added to shared-memory accumulators:
// the D x D exponentials are NOT recomputed by a separate parameter-gradient pass.
template <int D>
__device__ __forceinline__ void df_rows_k_bwd(const float* __restrict__ chunk, int n, int row0, const DfSmem& sm, const float (&x)[D],
                                              const float (&g)[D], float2 (&DX)[D], float (&dc)[D * D], int lane, int my_idx) {
  const float4* rows = reinterpret_cast<const float4*>(chunk);
#pragma unroll 1
  for (int m = 0; m < n; ++m) {
    float4 v[D];
#pragma unroll
    for (int i = 0; i < D; ++i) v[i] = rows[m * D + i];
    float2 d[D], p[D], tt[D], gs[D], r2 = make_float2(0.f, 0.f), WV = make_float2(0.f, 0.f);
    float red[4 * D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      d[k] = sub2(bc(x[k]), f2at(v, k));
      r2 = fma2(d[k], d[k], r2);
      p[k] = mul2(f2at(v, D + k), d[k]);
      tt[k] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const float2 u = mul2(bc(g[j]), d[j]);
      const float2 ur = mul2(u, r2), u2 = mul2(u, bc(kTwoLog2e));
      float2 s = make_float2(0.f, 0.f), wj = make_float2(0.f, 0.f), ejj, kjj;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float4 kc = lds_const4(sm.kc + i * D + j);
        const float2 e = ex2_2(fma2(r2, lo(kc), hi(kc)));
        const float2 pe = mul2(p[i], e);
        s = add2(s, pe);
        wj = fma2(lo(kc), pe, wj);
        tt[i] = fma2(u, e, tt[i]);
        const float2 ua = fma2(lo(kc), ur, u2);            // u_j (r2 k_ij + 2 log2e)
        dc[i * D + j] = fmaf(pe.x, ua.x, dc[i * D + j]);
        dc[i * D + j] = fmaf(pe.y, ua.y, dc[i * D + j]);
        if (i == j) {
          ejj = e;
          kjj = lo(kc);
        }
      }
      const float hj = sm.h[j];
      const float2 hr = sub2(bc(hj), r2);
      const float2 ge = mul2(bc(g[j]), ejj);
      const float2 gne = mul2(ge, f2at(v, D + j));
      // WV accumulates (W + V) / (-2 ln2):  W = -2 ln2 sum u_j wj ;  V = g_j nu_j e_jj (2 - 2 ln2 k_jj (h_j - r2))
      WV = fma2(u, wj, WV);
      WV = fma2(gne, fma2(kjj, hr, bc(-kLog2e)), WV);
      gs[j] = mul2(bc(g[j]), s);
      const float2 dnu_diag = mul2(ge, hr);                // g_j e_jj (h_j - r2)
      red[j] = dnu_diag.x;
      red[D + j] = dnu_diag.y;
      const float2 dg = mul2(gne, fma2(fma2(r2, kjj, bc(kTwoLog2e)), hr, bc(-hj * kLog2e)));
      dc[j * D + j] += dg.x + dg.y;
    }
    const float2 wv = mul2(WV, bc(2.f * kLn2));            // = -(W + V)
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float2 dnu = fma2(d[k], tt[k], make_float2(red[k], red[D + k]));
      float2 ddk = fma2(f2at(v, D + k), tt[k], gs[k]);
      ddk = fma2(d[k], wv, ddk);
      DX[k] = add2(DX[k], ddk);
      red[k] = dnu.x;
      red[D + k] = dnu.y;
      red[2 * D + k] = -ddk.x;
      red[3 * D + k] = -ddk.y;
    }
    const float tot = XRed<4 * D, 16>::run(red, lane);
    if (my_idx >= 0) atomicAdd(&sm.pacc[(row0 + m) * 4 * D + my_idx], tot);
  }
}

template <int D_, int R_, int W_ = 0>
struct DfPolicy {
  static constexpr int DP = D_;
  static constexpr int D = D_;
  static constexpr int R = R_;
  static constexpr int W = W_;              // 0: thread <-> R states; > 0: 32 states per CTA, W warps splitting the rows (small batches)
  static constexpr int kThreads = W_ ? 32 * W_ : 128;
  static constexpr int kMinBlocks = W_ ? 1 : 3;
  static constexpr int kStateThreads = W_ ? 32 : 0;
  static constexpr int kXsStride = W_ ? 32 : 0;       // 0: staging buffers xs / dx are strided by the block size
  static constexpr bool kCoopGlue = W_ > 0;           // sweep.cuh: solver glue spread over all threads
  static constexpr int kThreadsBwd = W_ ? 32 * W_ : 128;
  static constexpr int kMinBlocksBwd = W_ ? 1 : (D_ <= 6 ? 3 : 2);   // the reverse sweep also carries the D x D lengthscale accumulators
  static_assert(W_ == 0 || R_ == 1, "the warp-split instantiation runs one state per lane");
  using Geom = DfGeom;
  using Accum = DfAccum;
  using Smem = DfSmem;
  __device__ static __forceinline__ int xstride() { return W_ ? 32 : static_cast<int>(blockDim.x); }          // stride of the staging buffers
  __device__ static __forceinline__ int sslot() { return W_ ? static_cast<int>(threadIdx.x & 31) : static_cast<int>(threadIdx.x); }

  __device__ static __forceinline__ void finish(Smem&) {}
  __device__ static __forceinline__ Smem carve(float* smem, const Geom& g) {
    Smem s;
    s.bars = reinterpret_cast<uint64_t*>(smem);
    s.stages = smem + 32;
    float* hdr = s.stages + kPipeStages * g.stage_floats;
    s.kc = reinterpret_cast<const float4*>(hdr);
    s.h = hdr + 4 * D * D;
    s.xch = hdr + g.hdr_floats;
    s.ev = 0;
    s.loc = s.xch + (gridDim.z > 1 ? 2 * static_cast<int>(gridDim.z) * 2 * D * R * xstride() : 0);
    s.xs = s.loc + (W_ ? W_ * 2 * D * 32 : 0);
    s.dx = s.xs + D * R * xstride();
    s.dell = s.dx + D * R * xstride();
    s.dvar = s.dell + D * D;
    s.dcs = s.dvar + D;
    s.pacc = s.dcs + D * D;
    return s;
  }
  __device__ static __forceinline__ long setup(Smem& sm, ChunkPipe& pipe, const Geom& g, const float* packed, long n_evals, bool bwd) {
    const long total = n_evals * chunks_per_eval(g.cg);
    float* hdr = sm.stages + kPipeStages * g.stage_floats;
    for (int i = threadIdx.x; i < g.hdr_floats; i += blockDim.x) hdr[i] = packed[i];
    for (int i = threadIdx.x; i < D * R * xstride(); i += blockDim.x) sm.xs[i] = 0.f;
    if (bwd)
      for (int i = threadIdx.x; i < 2 * D * D + D + g.MP2 * 4 * D; i += blockDim.x) sm.dell[i] = 0.f;
    pipe.init(sm.stages, sm.bars, df_rows_ptr(packed, g, blockIdx.y), g.cg, total);  // contains the publishing __syncthreads
    return total;
  }
  __device__ static __forceinline__ void load_x(const Smem& sm, float (&x)[R][D]) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int d = 0; d < D; ++d) x[r][d] = sm.xs[(d * R + r) * xstride() + sslot()];
  }

  template <class Store>
  __device__ static __forceinline__ void eval_fwd(ChunkPipe& pipe, const Geom& g, long total, Smem& sm, Store&& store) {
    const int C = gridDim.z, rank = blockIdx.z;   // cluster = the z extent of the grid (1: no cluster)
    const int NS = C * (W_ ? W_ : 1), si = rank * (W_ ? W_ : 1) + (W_ ? static_cast<int>(threadIdx.x >> 5) : 0);   // row slices, this warp's
    if constexpr (W_ > 0) __syncthreads();   // the stage input staged by the state warp is visible to every warp
    float x[R][D];
    load_x(sm, x);
    float2 acc[R][D];
    float fpv[D][R];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) acc[r][c] = make_float2(0.f, 0.f);
    for (int a = 0; a < D; ++a)
      for (int c = 0; c < g.NCs; ++c) {
        const float* chunk = pipe.acquire(g.cg);
        int lo, cnt;
        df_slice(min(g.RCs, g.SP2 - c * g.RCs), NS, si, lo, cnt);
        df_rows_prior_fwd<D, R>(chunk + lo * g.rowf_s, cnt, x, acc);
        pipe.release(g.cg, total);
      }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        fpv[c][r] = acc[r][c].x + acc[r][c].y;
        acc[r][c] = make_float2(0.f, 0.f);
      }
    for (int c = 0; c < g.NCm; ++c) {
      const float* chunk = pipe.acquire(g.cg);
      int lo, cnt;
      df_slice(min(g.RCm, g.MP2 - c * g.RCm), NS, si, lo, cnt);
      df_rows_k_fwd<D, R>(chunk + lo * g.rowf_m, cnt, sm, x, acc);
      pipe.release(g.cg, total);
    }
    if constexpr (W_ > 0) {   // the warps' partial sums meet in shared memory; warp 0 carries the CTA's sums on
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        sm.loc[(warp * 2 * D + 2 * k) * 32 + lane] = fpv[k][0];
        sm.loc[(warp * 2 * D + 2 * k + 1) * 32 + lane] = acc[0][k].x + acc[0][k].y;
      }
      __syncthreads();
      if (warp == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
          float a = 0.f, b = 0.f;
          for (int w = 0; w < W_; ++w) {
            a += sm.loc[(w * 2 * D + 2 * k) * 32 + lane];
            b += sm.loc[(w * 2 * D + 2 * k + 1) * 32 + lane];
          }
          fpv[k][0] = a;
          acc[0][k] = make_float2(b, 0.f);
        }
      }
    }
    const bool owner = W_ == 0 || threadIdx.x < 32;   // threads that hold a state's sums
    if (C == 1) {
      if (owner) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
          float fu[R];
#pragma unroll
          for (int r = 0; r < R; ++r) fu[r] = acc[r][k].x + acc[r][k].y;
          store(k, fpv[k], fu);
        }
      }
    } else {   // all-gather the partial sums of the C row slices, then every CTA finishes all outputs
      const int T = xstride(), tid = sslot();
      float* xb = sm.xch + (sm.ev & 1) * C * 2 * D * R * T;
      if (owner) {
#pragma unroll
        for (int k = 0; k < D; ++k)
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const uint32_t a0 = smem_u32(xb + ((rank * 2 * D + 2 * k) * R + r) * T + tid);
            for (int q = 0; q < C; ++q) {
              df_st_cluster(a0, q, fpv[k][r]);
              df_st_cluster(a0 + R * T * 4, q, acc[r][k].x + acc[r][k].y);
            }
          }
      }
      df_cluster_sync();
      if (owner) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
          float fp[R], fu[R];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            fp[r] = fu[r] = 0.f;
            for (int q = 0; q < C; ++q) {
              fp[r] += xb[((q * 2 * D + 2 * k) * R + r) * T + tid];
              fu[r] += xb[((q * 2 * D + 2 * k + 1) * R + r) * T + tid];
            }
          }
          store(k, fp, fu);
        }
      }
    }
    ++sm.ev;
  }

  // all-output VJP at one evaluation: leaves dL/dx in sm.dx, folds the lengthscale (theta path) and variance statistics
  __device__ static __forceinline__ void vjp(ChunkPipe& pipe, const Geom& g, long total, Smem& sm, const States<R>& st,
                                             const float* gvec, const float* fvec, const float* fpvec, long kstride, long sstride) {
    const int lane = threadIdx.x & 31;
    const int C = gridDim.z, rank = blockIdx.z;   // cluster = the z extent of the grid (1: no cluster)
    const int NS = C * (W_ ? W_ : 1), si = rank * (W_ ? W_ : 1) + (W_ ? static_cast<int>(threadIdx.x >> 5) : 0);   // row slices, this warp's
    if constexpr (W_ > 0) __syncthreads();   // the stage input staged by the state warp is visible to every warp
    float x[R][D], gg[R][D], dxs[R][D];
    load_x(sm, x);
    long s_of[R];
    bool ok_of[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if constexpr (W_ > 0) {   // every warp of the CTA works on the same 32 states: lane <-> state
        const long n = static_cast<long>(blockIdx.x) * 32 + lane;
        ok_of[r] = n < g.N;
        s_of[r] = static_cast<long>(blockIdx.y) * g.N + (ok_of[r] ? n : g.N - 1);
      } else {
        ok_of[r] = st.ok[r];
        s_of[r] = st.s[r];
      }
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const long at = k * kstride + s_of[r] * sstride;
        gg[r][k] = ok_of[r] ? gvec[at] : 0.f;
        v += gg[r][k] * (fvec[at] - 0.5f * fpvec[at]);
        dxs[r][k] = 0.f;
      }
      v = warp_sum(v);
      if (lane == 0 && si == 0) atomicAdd(&sm.dvar[k], v);   // (no sum over rows in it: one warp of the cluster takes it)
    }
    for (int a = 0; a < D; ++a) {
      float Q[R][D];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int d = 0; d < D; ++d) Q[r][d] = 0.f;
      for (int c = 0; c < g.NCs; ++c) {
        const float* chunk = pipe.acquire(g.cg);
        int lo, cnt;
        df_slice(min(g.RCs, g.SP2 - c * g.RCs), NS, si, lo, cnt);
        df_rows_prior_bwd<D, R>(chunk + lo * g.rowf_s, cnt, x, gg, Q);
        pipe.release(g.cg, total);
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
        float u = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          dxs[r][d] += Q[r][d];
          u = fmaf(x[r][d], Q[r][d], u);   // padded lanes carry g = 0 -> Q = 0
        }
        u = warp_sum(u);
        if (lane == 0) atomicAdd(&sm.dell[a * D + d], u);
      }
    }
    static_assert(R == 1, "the DF reverse sweep runs one state per thread");
    float2 DX[D];
    float dc[D * D];
#pragma unroll
    for (int d = 0; d < D; ++d) DX[d] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < D * D; ++i) dc[i] = 0.f;
    const int my_idx = xred_index(4 * D, lane);
    for (int c = 0; c < g.NCm; ++c) {
      const float* chunk = pipe.acquire(g.cg);
      int lo, cnt;
      df_slice(min(g.RCm, g.MP2 - c * g.RCm), NS, si, lo, cnt);
      df_rows_k_bwd<D>(chunk + lo * g.rowf_m, cnt, c * g.RCm + lo, sm, x[0], gg[0], DX, dc, lane, my_idx);
      pipe.release(g.cg, total);
    }
#pragma unroll
    for (int i = 0; i < D * D; ++i) {
      const float v = warp_sum(dc[i]);
      if (lane == 0) atomicAdd(&sm.dcs[i], v);
    }
    float part[D];
#pragma unroll
    for (int d = 0; d < D; ++d) part[d] = dxs[0][d] + DX[d].x + DX[d].y;
    if constexpr (W_ > 0) {   // the warps' partial J^T g meet in shared memory; warp 0 carries the CTA's sum on
      const int warp = threadIdx.x >> 5;
#pragma unroll
      for (int d = 0; d < D; ++d) sm.loc[(warp * 2 * D + d) * 32 + lane] = part[d];
      __syncthreads();
      if (warp == 0) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          float v = 0.f;
          for (int w = 0; w < W_; ++w) v += sm.loc[(w * 2 * D + d) * 32 + lane];
          part[d] = v;
        }
      }
    }
    const bool owner = W_ == 0 || threadIdx.x < 32;
    const int T = xstride(), tid = sslot();
    if (C == 1) {
      if (owner) {
#pragma unroll
        for (int d = 0; d < D; ++d) sm.dx[d * T + tid] = part[d];
      }
    } else {   // all-gather the partial J^T g of the C row slices
      float* xb = sm.xch + (sm.ev & 1) * C * 2 * D * T;
      if (owner) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const uint32_t a0 = smem_u32(xb + (rank * 2 * D + d) * T + tid);
          for (int q = 0; q < C; ++q) df_st_cluster(a0, q, part[d]);
        }
      }
      df_cluster_sync();
      if (owner) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          float v = 0.f;
          for (int q = 0; q < C; ++q) v += xb[(q * 2 * D + d) * T + tid];
          sm.dx[d * T + tid] = v;
        }
      }
    }
    ++sm.ev;
    if constexpr (W_ > 0) __syncthreads();   // dx is consumed by the cooperative glue (all threads); loc may be rewritten
  }

  __device__ static __forceinline__ void flush(const Smem& sm, const Geom& g, const Accum& acc) {
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
      atomicAdd(&acc.dell_x[i], sm.dell[i]);
      atomicAdd(&acc.dc[i], sm.dcs[i]);
    }
    for (int i = threadIdx.x; i < D; i += blockDim.x) atomicAdd(&acc.dvar[i], sm.dvar[i]);
    const int l = blockIdx.y;
    for (int i = threadIdx.x; i < g.MP2 * 4 * D; i += blockDim.x) {
      const int pair = i / (4 * D), q = i - pair * 4 * D;
      const int what = q / (2 * D), h = (q / D) & 1, k = q % D;
      const float v = sm.pacc[i];
      if (what == 0) atomicAdd(&acc.dnu[(static_cast<size_t>(l) * 2 * g.MP2 + 2 * pair + h) * D + k], v);
      else atomicAdd(&acc.dz[(2 * static_cast<size_t>(pair) + h) * D + k], v);
    }
  }
};

// =============================================================================================
// gradient of the feature operator: dB'[s,a,c] = sum_n g_c cos(theta'_sa)  (the inducing-point gradients dnu, dZ and the
// lengthscale statistics come out of the reverse sweep itself).  Threads <-> feature-pair rows (a, s); state evaluations
// (x, g) of one sample stream through shared memory in batches.
// =============================================================================================
constexpr int kDfPgBatch = 128;
constexpr int kDfPgThreads = 128;

template <int D>
__global__ void __launch_bounds__(kDfPgThreads, 4) k_df_pgrad(const DfPgradArgs a) {
  const DfGeom& g = a.g;
  constexpr int SROW = ((2 * D + 3) / 4) * 4;   // staged evaluation: x[D], g[D] (+pad)
  constexpr int NV = SROW / 4;
  __shared__ __align__(16) float stage[kDfPgBatch * SROW];
  const int l = blockIdx.z;
  const float* rows = df_rows_ptr(a.packed, g, l);
  float2 prm[D + 1];       // Om'[D], b'
  const int row = blockIdx.y * blockDim.x + threadIdx.x;   // a * SP2 + s-pair
  const bool active = row < g.D * g.SP2;
  if (active) {
    const float2* src = reinterpret_cast<const float2*>(rows + static_cast<size_t>(row) * g.rowf_s);
#pragma unroll
    for (int i = 0; i < D + 1; ++i) prm[i] = src[i];
  }
  float2 acc[D];
#pragma unroll
  for (int i = 0; i < D; ++i) acc[i] = make_float2(0.f, 0.f);

  const long total = a.n_te * g.N;
  const long per = (total + a.chunks_b - 1) / a.chunks_b;
  const long e_lo = static_cast<long>(blockIdx.x) * per;
  const long e_hi = e_lo + per < total ? e_lo + per : total;
  for (long e0 = e_lo; e0 < e_hi; e0 += kDfPgBatch) {
    for (int idx = threadIdx.x; idx < kDfPgBatch; idx += blockDim.x) {
      const long e = e0 + idx;
      float* srow = stage + idx * SROW;
      if (e < e_hi) {
        const long te = e / g.N;
        const long s = static_cast<long>(l) * g.N + (e - te * g.N);
#pragma unroll
        for (int d = 0; d < D; ++d) {
          srow[d] = a.xsave[(te * D + d) * g.NL + s];
          srow[D + d] = a.gsave[(te * D + d) * g.NL + s];
        }
      }
    }
    __syncthreads();
    if (active) {
      const int nb = (e_hi - e0) < kDfPgBatch ? static_cast<int>(e_hi - e0) : kDfPgBatch;
#pragma unroll 2
      for (int idx = 0; idx < nb; ++idx) {
        const float4* srow = reinterpret_cast<const float4*>(stage + idx * SROW);
        float xv[SROW];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float4 q = srow[i];
          xv[4 * i] = q.x; xv[4 * i + 1] = q.y; xv[4 * i + 2] = q.z; xv[4 * i + 3] = q.w;
        }
        float2 th = prm[D];
#pragma unroll
        for (int k = 0; k < D; ++k) th = fma2(bc(xv[k]), prm[k], th);
        const float2 cs = cos_2(th);
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = fma2(bc(xv[D + c]), cs, acc[c]);
      }
    }
    __syncthreads();
  }
  if (active) {
    const size_t base = (static_cast<size_t>(l) * g.D * g.SP2 + row) * 2 * D;   // [l][a][s-pair][even/odd][c]
#pragma unroll
    for (int c = 0; c < D; ++c) {
      atomicAdd(&a.acc.dbp[base + c], acc[c].x);
      atomicAdd(&a.acc.dbp[base + D + c], acc[c].y);
    }
  }
}

// launch-shape heuristic of the DF sweep kernels (states per CTA = 128 * R)
inline void df_pick_shape(const DfGeom& g, bool bwd, int& threads, int& R) {
  const long want = 2L * 148;
  const int rmax = (!bwd && g.D <= 6) ? 2 : 1;
  const int cand[5][2] = {{128, rmax}, {128, 1}, {64, 1}, {32, 1}, {32, 1}};
  for (int i = 0; i < 5; ++i) {
    threads = cand[i][0];
    R = cand[i][1];
    const long per = static_cast<long>(threads) * R;
    if (((static_cast<long>(g.N) + per - 1) / per) * g.L >= want) return;
  }
}

}  // namespace gpode
