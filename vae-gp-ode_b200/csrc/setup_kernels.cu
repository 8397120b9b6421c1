// Per-rollout setup of the sparse GP at the M inducing points, batched over output dimensions and MC samples
// (sm_100a): the pieces of SVGP_Layer.build_cache / compute_nu / kl that sit in the autograd graph
// (reference experiments/model/core/svpy.py:88-121,144-175; core/kernels.py:98-110,155-172).
//
//   nu = Lc^-T (u - Lc^-1 u_prior),  Lc = chol(K(Z,Z) + 1e-5 I)        one (M x M) system per output dim (dimwise)
//                                                                       or one shared by all output dims (shared)
// Layout used inside: right-hand sides are (Kc, M, NR) row-major, Kc = number of matrices, NR = columns per
// matrix (dimwise: the L samples; shared: L * D_out).  Matrices are (Kc, M, M) row-major; after the Cholesky the
// lower triangle holds Lc and the strict upper triangle is garbage (never read).
//
// Backward (closed form, derivation checked against autograd in tests/test_setup_algebra.py): with
//   bb = Lc^-1 nu_bar, r = Lc^-T bb, a = Lc^-1 u_prior (saved), q = Lc^-T a,
//   S_ij = -1/2 sum_cols u[max(i,j)] bb[min(i,j)]
//   A_bar = Lc^-T S Lc^-1 + 1/2 sum_cols (r q^T + q r^T),   u_bar = bb,   u_prior_bar = -r
// and A_bar is contracted with dK/d(Z, ell, var) on the fly (A_bar is never symmetrised explicitly: it is symmetric).
#include "common.cuh"
#include "setup.h"

namespace gpode {

namespace {

constexpr int NB = 32;            // block size of the blocked factorisation / substitutions
constexpr int kSetupThreads = 256;
// The factorisation, the substitutions and the contraction of the backward run in double precision: K(Z,Z) + jitter has
// cond 1e4..1e6 at the reference's settings and an fp32 Cholesky is the largest single error of the reference's own
// nu and leaf gradients; the systems are tiny (M <= 512) next to the rollout, B200 has full-rate FP64 FMA units, and inputs / outputs
// stay fp32 at the ABI.  (tests/test_gpu_setup.py: nu and its gradients against the fp64 oracle.)
using real = double;

// element (k, i, r) of a reference-layout (L, M, D_out) tensor seen as the (Kc, M, NR) right-hand-side array
__device__ __forceinline__ size_t lmd_index(const NuGeom& g, int k, int i, int r) {
  const int l = g.dimwise ? r : r / g.D_out;
  const int kk = g.dimwise ? k : r - l * g.D_out;
  return (static_cast<size_t>(l) * g.M + i) * g.D_out + kk;
}
// element (k, i, r) of nu in its reference layout: dimwise (L, D_out, M, 1); shared (L, M, D_out)
__device__ __forceinline__ size_t nu_index(const NuGeom& g, int k, int i, int r) {
  if (g.dimwise) return (static_cast<size_t>(r) * g.D_out + k) * g.M + i;
  return lmd_index(g, k, i, r);
}
__device__ __forceinline__ real warp_sum_real(real v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float ell_of(const NuGeom& g, const float* ell, int k, int d) { return g.dimwise ? ell[k * g.D_in + d] : ell[d]; }

// ---------------------------------------------------------------------------------------------
// K(Z,Z) + jitter I  (core/kernels.py:98-110 evaluated with direct differences)
// ---------------------------------------------------------------------------------------------
__global__ void k_kzz_build(const NuGeom g, const float* __restrict__ Z, const float* __restrict__ ell, const float* __restrict__ var,
                            real* __restrict__ A) {
  const int k = blockIdx.y;
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long>(g.M) * g.M) return;
  const int i = static_cast<int>(idx / g.M), j = static_cast<int>(idx - static_cast<long>(i) * g.M);
  real sq = 0;
  for (int d = 0; d < g.D_in; ++d) {
    const real t = (static_cast<real>(Z[i * g.D_in + d]) - static_cast<real>(Z[j * g.D_in + d])) / static_cast<real>(ell_of(g, ell, k, d));
    sq = fma(t, t, sq);
  }
  real v = static_cast<real>(var[g.dimwise ? k : 0]) * exp(-0.5 * sq);
  if (i == j) v += static_cast<real>(g.jitter);
  A[(static_cast<size_t>(k) * g.M + i) * g.M + j] = v;
}

// ---------------------------------------------------------------------------------------------
// blocked right-looking Cholesky, one CTA per matrix, matrix in global memory (L2 resident: <= 1 MB),
// the current 32-column panel transposed in shared memory for the trailing update.
// info[k] = 1 + index of the first non-positive pivot (0 = ok), like LAPACK potrf.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSetupThreads) k_chol(const int M, real* __restrict__ Aall, int* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  real* sm = reinterpret_cast<real*>(sm_raw);
  real* Dg = sm;                       // [NB][NB+1] diagonal block
  real* Pt = sm + NB * (NB + 1);       // [NB][Mp] panel, transposed: Pt[c][row - row0]
  const int Mp = (M + NB - 1) / NB * NB;
  real* A = Aall + static_cast<size_t>(blockIdx.x) * M * M;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;
  for (int j0 = 0; j0 < M; j0 += NB) {
    const int nbk = min(NB, M - j0);
    // (a) diagonal block -> shared, factor with warp 0 (rows >= nbk act as identity)
    for (int e = tid; e < NB * NB; e += blockDim.x) {
      const int r = e / NB, c = e - r * NB;
      Dg[r * (NB + 1) + c] = (r < nbk && c < nbk) ? A[static_cast<size_t>(j0 + r) * M + j0 + c] : (r == c ? real(1) : real(0));
    }
    __syncthreads();
    if (warp == 0) {
      for (int c = 0; c < NB; ++c) {
        real piv = Dg[c * (NB + 1) + c];
        if (!(piv > real(0)) && c < nbk && lane == 0 && s_bad == 0) s_bad = j0 + c + 1;
        piv = sqrt(piv);
        __syncwarp();
        if (lane == c) Dg[c * (NB + 1) + c] = piv;
        if (lane > c) Dg[lane * (NB + 1) + c] /= piv;
        __syncwarp();
        if (lane > c) {
          const real lrc = Dg[lane * (NB + 1) + c];
          for (int cc = c + 1; cc <= lane; ++cc) Dg[lane * (NB + 1) + cc] -= lrc * Dg[cc * (NB + 1) + c];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    for (int e = tid; e < nbk * nbk; e += blockDim.x) {
      const int r = e / nbk, c = e - r * nbk;
      if (c <= r) A[static_cast<size_t>(j0 + r) * M + j0 + c] = Dg[r * (NB + 1) + c];
    }
    const int r0 = j0 + NB;             // first row below the diagonal block
    const int nrows = M - r0;
    if (nrows <= 0) break;
    // (b) panel: X Lbb^T = A[r0:, j0:j0+nb]  -> one thread per row, forward substitution over the 32 columns
    for (int rr = tid; rr < nrows; rr += blockDim.x) {
      real x[NB];
      real* arow = A + static_cast<size_t>(r0 + rr) * M + j0;
#pragma unroll
      for (int c = 0; c < NB; ++c) x[c] = c < nbk ? arow[c] : real(0);
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        real v = x[c];
#pragma unroll
        for (int cc = 0; cc < c; ++cc) v = fma(-x[cc], Dg[c * (NB + 1) + cc], v);
        x[c] = v / Dg[c * (NB + 1) + c];
      }
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        if (c < nbk) arow[c] = x[c];
        Pt[c * Mp + rr] = x[c];
      }
    }
    __syncthreads();
    // (c) trailing update of the lower triangle: A[i][j] -= sum_c P[i][c] P[j][c], 4 x 4 register tiles
    const int nt = (nrows + 3) / 4;
    for (long t = tid; t < static_cast<long>(nt) * nt; t += blockDim.x) {
      const int ti = static_cast<int>(t / nt), tj = static_cast<int>(t - static_cast<long>(ti) * nt);
      if (tj > ti) continue;
      real acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = real(0);
#pragma unroll 4
      for (int c = 0; c < NB; ++c) {
        const double4 pi = *reinterpret_cast<const double4*>(Pt + c * Mp + 4 * ti);
        const double4 pj = *reinterpret_cast<const double4*>(Pt + c * Mp + 4 * tj);
        const real vi[4] = {pi.x, pi.y, pi.z, pi.w}, vj[4] = {pj.x, pj.y, pj.z, pj.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fma(vi[a], vj[b], acc[a][b]);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = 4 * ti + a;
        if (i >= nrows) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int j = 4 * tj + b;
          if (j <= i) A[static_cast<size_t>(r0 + i) * M + r0 + j] -= acc[a][b];
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0 && info) info[blockIdx.x] = s_bad;
}

// ---------------------------------------------------------------------------------------------
// blocked triangular solve with up to 32 right-hand-side columns per CTA, RHS block resident in shared memory:
//   TRANS = false:  Lc X = B  (top to bottom)        TRANS = true:  Lc^T X = B  (bottom to top)
// src / dst are (Kc, M, NR) arrays (dst may alias src); src_t: read the source transposed ((Kc, NR, M), NR == M).
// ---------------------------------------------------------------------------------------------
template <bool TRANS>
__global__ void __launch_bounds__(kSetupThreads) k_trsm(const int M, const int NR, const real* __restrict__ Lall, const real* src,
                                                        real* dst, const int src_t) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  real* sm = reinterpret_cast<real*>(sm_raw);
  const int Mp = (M + NB - 1) / NB * NB;
  real* Xs = sm;                          // [Mp][NB+1]
  real* Lb = Xs + Mp * (NB + 1);          // [NB][NB+1] diagonal block
  real* Ls = Lb + NB * (NB + 1);          // [64][NB+1] rows of the off-diagonal strip
  const int k = blockIdx.y, c0 = blockIdx.x * NB;
  const int ncol = min(NB, NR - c0);
  const real* Lc = Lall + static_cast<size_t>(k) * M * M;
  const real* S = src + static_cast<size_t>(k) * M * NR;
  real* Dd = dst + static_cast<size_t>(k) * M * NR;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  if (src_t) {
    for (int e = tid; e < Mp * NB; e += blockDim.x) {
      const int c = e / Mp, i = e - c * Mp;                        // consecutive threads walk i: coalesced rows of the source
      Xs[i * (NB + 1) + c] = (i < M && c < ncol) ? S[static_cast<size_t>(c0 + c) * M + i] : real(0);
    }
  } else {
    for (int e = tid; e < Mp * NB; e += blockDim.x) {
      const int i = e / NB, c = e - i * NB;
      Xs[i * (NB + 1) + c] = (i < M && c < ncol) ? S[static_cast<size_t>(i) * NR + c0 + c] : real(0);
    }
  }
  const int nblk = Mp / NB;
  for (int bi = 0; bi < nblk; ++bi) {
    const int b = TRANS ? nblk - 1 - bi : bi;
    const int j0 = b * NB;
    __syncthreads();
    for (int e = tid; e < NB * NB; e += blockDim.x) {
      const int r = e / NB, c = e - r * NB;
      const int gr = j0 + r, gc = j0 + c;
      Lb[r * (NB + 1) + c] = (gr < M && gc < M && gc <= gr) ? Lc[static_cast<size_t>(gr) * M + gc] : (r == c ? real(1) : real(0));
    }
    __syncthreads();
    if (warp == 0) {                       // lane = column: 32-step substitution on the diagonal block
      real x[NB];
#pragma unroll
      for (int r = 0; r < NB; ++r) x[r] = Xs[(j0 + r) * (NB + 1) + lane];
      if (!TRANS) {
#pragma unroll
        for (int r = 0; r < NB; ++r) {
          real v = x[r];
#pragma unroll
          for (int c = 0; c < r; ++c) v = fma(-Lb[r * (NB + 1) + c], x[c], v);
          x[r] = v / Lb[r * (NB + 1) + r];
        }
      } else {
#pragma unroll
        for (int r = NB - 1; r >= 0; --r) {
          real v = x[r];
#pragma unroll
          for (int c = r + 1; c < NB; ++c) v = fma(-Lb[c * (NB + 1) + r], x[c], v);
          x[r] = v / Lb[r * (NB + 1) + r];
        }
      }
#pragma unroll
      for (int r = 0; r < NB; ++r) Xs[(j0 + r) * (NB + 1) + lane] = x[r];
    }
    __syncthreads();
    // update the remaining rows: non-trans rows i >= j0 + NB use Lc[i][j0 + c]; trans rows j < j0 use Lc[j0 + c][j]
    const int lo_row = TRANS ? 0 : j0 + NB, hi_row = TRANS ? j0 : Mp;
    for (int s0 = lo_row; s0 < hi_row; s0 += 64) {
      const int ns = min(64, hi_row - s0);
      if (!TRANS) {
        for (int e = tid; e < ns * NB; e += blockDim.x) {
          const int rr = e / NB, c = e - rr * NB;
          const int gi = s0 + rr, gc = j0 + c;
          Ls[rr * (NB + 1) + c] = (gi < M && gc < M) ? Lc[static_cast<size_t>(gi) * M + gc] : real(0);
        }
      } else {
        for (int e = tid; e < ns * NB; e += blockDim.x) {
          const int c = e / ns, rr = e - c * ns;                    // consecutive threads walk the row of Lc: coalesced
          const int gj = s0 + rr, gc = j0 + c;
          Ls[rr * (NB + 1) + c] = (gc < M && gj < M) ? Lc[static_cast<size_t>(gc) * M + gj] : real(0);
        }
      }
      __syncthreads();
      for (int rr = warp; rr < ns; rr += nwarp) {
        real acc = 0;
#pragma unroll
        for (int c = 0; c < NB; ++c) acc = fma(Ls[rr * (NB + 1) + c], Xs[(j0 + c) * (NB + 1) + lane], acc);
        Xs[(s0 + rr) * (NB + 1) + lane] -= acc;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  for (int e = tid; e < M * NB; e += blockDim.x) {
    const int i = e / NB, c = e - i * NB;
    if (c < ncol) Dd[static_cast<size_t>(i) * NR + c0 + c] = Xs[i * (NB + 1) + c];
  }
}

// ---------------------------------------------------------------------------------------------
// small glue kernels (all elementwise over (k, i, r) or (k, i, j))
// ---------------------------------------------------------------------------------------------
// rhs[k][i][r] = reference-layout (L, M, D_out) tensor
__global__ void k_gather_lmd(const NuGeom g, const float* __restrict__ src, real* __restrict__ rhs) {
  const long n = static_cast<long>(g.Kc) * g.M * g.NR;
  for (long e = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += static_cast<long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(e % g.NR), i = static_cast<int>((e / g.NR) % g.M), k = static_cast<int>(e / (static_cast<long>(g.NR) * g.M));
    rhs[e] = src[lmd_index(g, k, i, r)];
  }
}
// b = u - a  (u in reference layout), in place on the a array
__global__ void k_u_minus_a(const NuGeom g, const float* __restrict__ u, real* __restrict__ a_then_b) {
  const long n = static_cast<long>(g.Kc) * g.M * g.NR;
  for (long e = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += static_cast<long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(e % g.NR), i = static_cast<int>((e / g.NR) % g.M), k = static_cast<int>(e / (static_cast<long>(g.NR) * g.M));
    a_then_b[e] = static_cast<real>(u[lmd_index(g, k, i, r)]) - a_then_b[e];
  }
}
// scatter (Kc, M, NR) -> nu layout (mode 0) or (L, M, D_out) layout with a sign (mode 1)
__global__ void k_scatter(const NuGeom g, const real* __restrict__ rhs, float* __restrict__ out, const int mode, const float scale) {
  const long n = static_cast<long>(g.Kc) * g.M * g.NR;
  for (long e = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += static_cast<long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(e % g.NR), i = static_cast<int>((e / g.NR) % g.M), k = static_cast<int>(e / (static_cast<long>(g.NR) * g.M));
    out[mode == 0 ? nu_index(g, k, i, r) : lmd_index(g, k, i, r)] = static_cast<float>(scale * rhs[e]);
  }
}
// gather nu-layout tensor into (Kc, M, NR)
__global__ void k_gather_nu(const NuGeom g, const float* __restrict__ src, real* __restrict__ rhs) {
  const long n = static_cast<long>(g.Kc) * g.M * g.NR;
  for (long e = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += static_cast<long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(e % g.NR), i = static_cast<int>((e / g.NR) % g.M), k = static_cast<int>(e / (static_cast<long>(g.NR) * g.M));
    rhs[e] = src[nu_index(g, k, i, r)];
  }
}
// S[k][i][j] = -1/2 sum_r u[k][max(i,j)][r] bb[k][min(i,j)][r]
__global__ void k_build_S(const NuGeom g, const float* __restrict__ u, const real* __restrict__ bb, real* __restrict__ S) {
  const int k = blockIdx.y;
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long>(g.M) * g.M) return;
  const int i = static_cast<int>(idx / g.M), j = static_cast<int>(idx - static_cast<long>(i) * g.M);
  const int hi_ = i > j ? i : j, lo_ = i > j ? j : i;
  real acc = 0;
  for (int r = 0; r < g.NR; ++r) acc = fma(static_cast<real>(u[lmd_index(g, k, hi_, r)]), bb[(static_cast<size_t>(k) * g.M + lo_) * g.NR + r], acc);
  S[(static_cast<size_t>(k) * g.M + i) * g.M + j] = -0.5 * acc;
}
// A_bar = X + 1/2 sum_r (r_i q_j + q_i r_j) contracted with dK/d(var, ell, Z); one warp per (k, i) row
__global__ void k_kzz_bwd(const NuGeom g, const float* __restrict__ Z, const float* __restrict__ ell, const float* __restrict__ var,
                          const real* __restrict__ X, const real* __restrict__ rr_, const real* __restrict__ qq_, float* __restrict__ d_Z,
                          float* __restrict__ d_ell, float* __restrict__ d_var) {
  const int k = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= g.M) return;
  const real vk = var[g.dimwise ? k : 0];
  real zi[kMaxD], inv2[kMaxD], dz[kMaxD], dl[kMaxD], dv = 0;
  for (int d = 0; d < g.D_in; ++d) {
    zi[d] = Z[i * g.D_in + d];
    const real e = ell_of(g, ell, k, d);
    inv2[d] = 1.0 / (e * e);
    dz[d] = 0;
    dl[d] = 0;
  }
  const real* ri = rr_ + (static_cast<size_t>(k) * g.M + i) * g.NR;
  const real* qi = qq_ + (static_cast<size_t>(k) * g.M + i) * g.NR;
  for (int j = lane; j < g.M; j += 32) {
    real ab = X[(static_cast<size_t>(k) * g.M + i) * g.M + j];
    const real* rj = rr_ + (static_cast<size_t>(k) * g.M + j) * g.NR;
    const real* qj = qq_ + (static_cast<size_t>(k) * g.M + j) * g.NR;
    real rk = 0;
    for (int r = 0; r < g.NR; ++r) rk += ri[r] * qj[r] + qi[r] * rj[r];
    ab = fma(0.5, rk, ab);
    real diff[kMaxD], sq = 0;
    for (int d = 0; d < g.D_in; ++d) {
      diff[d] = zi[d] - static_cast<real>(Z[j * g.D_in + d]);
      sq = fma(diff[d] * diff[d], inv2[d], sq);
    }
    const real E = exp(-0.5 * sq);
    const real G = ab * vk * E;                 // A_bar_ij K_ij
    dv = fma(ab, E, dv);                        // dK/dvar = E
    for (int d = 0; d < g.D_in; ++d) {
      dl[d] = fma(G * diff[d] * diff[d], inv2[d], dl[d]);      // x 1/ell below
      dz[d] = fma(-2.0 * G * diff[d], inv2[d], dz[d]);         // both (i,j) and (j,i) entries depend on Z_i (A_bar symmetric)
    }
  }
  dv = warp_sum_real(dv);
  if (lane == 0 && d_var) atomicAdd(&d_var[g.dimwise ? k : 0], static_cast<float>(dv));
  for (int d = 0; d < g.D_in; ++d) {
    const real a = warp_sum_real(dl[d]), b = warp_sum_real(dz[d]);
    if (lane == 0) {
      if (d_ell) atomicAdd(&d_ell[g.dimwise ? k * g.D_in + d : d], static_cast<float>(a / static_cast<real>(ell_of(g, ell, k, d))));
      if (d_Z) atomicAdd(&d_Z[i * g.D_in + d], static_cast<float>(b));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// inducing sample and whitened KL on the PACKED lower-triangular parameter (row-major tril order,
// misc/transforms.py:71-77): no (D_out, M, M) scatter is ever materialised.
//   u[l][n][d] = sum_{m<=n} Lq_d[n][m] eps[l][m][d] + Um[n][d]                       (svpy.py:88-101)
//   kl = 1/2 sum_d ( -sum_i log Lq_d[i][i]^2 + |Um[:,d]|^2 + |Lq_d|_F^2 - M )            (svpy.py:144-175)
// ---------------------------------------------------------------------------------------------
__global__ void k_inducing_fwd(const int L, const int M, const int D, const float* __restrict__ Lq, const float* __restrict__ Um,
                               const float* __restrict__ eps, float* __restrict__ u) {
  // one warp per (n, d); lanes split m
  const int lane = threadIdx.x & 31;
  const long wid = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (wid >= static_cast<long>(M) * D) return;
  const int n = static_cast<int>(wid / D), d = static_cast<int>(wid - static_cast<long>(n) * D);
  const float* row = Lq + static_cast<size_t>(d) * (static_cast<size_t>(M) * (M + 1) / 2) + static_cast<size_t>(n) * (n + 1) / 2;
  for (int l = 0; l < L; ++l) {
    float acc = 0.f;
    for (int m = lane; m <= n; m += 32) acc = fmaf(row[m], eps[(static_cast<size_t>(l) * M + m) * D + d], acc);
    acc = warp_sum(acc);
    if (lane == 0) u[(static_cast<size_t>(l) * M + n) * D + d] = acc + Um[n * D + d];
  }
}
// dLq_d[n][m] (+)= sum_l du[l][n][d] eps[l][m][d] ; dUm[n][d] (+)= sum_l du[l][n][d]     (accumulate = 1: add into the outputs)
__global__ void k_inducing_bwd(const int L, const int M, const int D, const float* __restrict__ eps, const float* __restrict__ du,
                               float* __restrict__ dLq, float* __restrict__ dUm, const int accumulate) {
  const size_t P = static_cast<size_t>(M) * (M + 1) / 2;
  const size_t total = P * D;
  for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(e / P);
    const size_t t = e - static_cast<size_t>(d) * P;
    int n = static_cast<int>((sqrtf(8.f * static_cast<float>(t) + 1.f) - 1.f) * 0.5f);
    while (static_cast<size_t>(n) * (n + 1) / 2 > t) --n;
    while (static_cast<size_t>(n + 1) * (n + 2) / 2 <= t) ++n;
    const int m = static_cast<int>(t - static_cast<size_t>(n) * (n + 1) / 2);
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc = fmaf(du[(static_cast<size_t>(l) * M + n) * D + d], eps[(static_cast<size_t>(l) * M + m) * D + d], acc);
    dLq[e] = accumulate ? dLq[e] + acc : acc;
    if (m == 0) {
      float s = 0.f;
      for (int l = 0; l < L; ++l) s += du[(static_cast<size_t>(l) * M + n) * D + d];
      dUm[n * D + d] = accumulate ? dUm[n * D + d] + s : s;
    }
  }
}
__global__ void k_kl_fwd(const int M, const int D, const float* __restrict__ Lq, const float* __restrict__ Um, float* __restrict__ kl) {
  const size_t P = static_cast<size_t>(M) * (M + 1) / 2;
  float acc = 0.f;
  for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < P * D; e += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = Lq[e];
    acc = fmaf(v, v, acc);
  }
  for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < static_cast<size_t>(M) * D; e += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(e / D), d = static_cast<int>(e - static_cast<size_t>(i) * D);
    const float m = Um[e], dg = Lq[static_cast<size_t>(d) * P + static_cast<size_t>(i) * (i + 1) / 2 + i];
    acc += m * m - logf(dg * dg) - 1.f;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(kl, 0.5f * acc);
}
__global__ void k_kl_bwd(const int M, const int D, const float* __restrict__ Lq, const float* __restrict__ Um, const float* __restrict__ dkl,
                         float* __restrict__ dLq, float* __restrict__ dUm) {
  const size_t P = static_cast<size_t>(M) * (M + 1) / 2;
  const float gk = dkl[0];
  for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < P * D; e += static_cast<size_t>(gridDim.x) * blockDim.x) dLq[e] = gk * Lq[e];
  for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < static_cast<size_t>(M) * D; e += static_cast<size_t>(gridDim.x) * blockDim.x) dUm[e] = gk * Um[e];
}
// diagonal part of the KL gradient, after k_kl_bwd: dLq[diag] -= dkl / Lq[diag]
__global__ void k_kl_bwd_diag(const int M, const int D, const float* __restrict__ Lq, const float* __restrict__ dkl, float* __restrict__ dLq) {
  const size_t P = static_cast<size_t>(M) * (M + 1) / 2;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M * D) return;
  const int i = e / D, d = e - i * D;
  const size_t at = static_cast<size_t>(d) * P + static_cast<size_t>(i) * (i + 1) / 2 + i;
  dLq[at] -= dkl[0] / Lq[at];
}

inline int blocks_for(long n, int threads, int cap = 148 * 8) {
  long b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}
inline size_t trsm_smem(int M) {
  const int Mp = (M + NB - 1) / NB * NB;
  return (static_cast<size_t>(Mp) * (NB + 1) + NB * (NB + 1) + 64 * (NB + 1)) * sizeof(real);
}
inline size_t chol_smem(int M) {
  const int Mp = (M + NB - 1) / NB * NB;
  return (static_cast<size_t>(NB) * (NB + 1) + static_cast<size_t>(NB) * Mp) * sizeof(real);
}

template <bool TRANS>
cudaError_t trsm(const NuGeom& g, int NR, const real* Lc, const real* src, real* dst, int src_t, cudaStream_t st) {
  const size_t smem = trsm_smem(g.M);
  cudaError_t e = cudaFuncSetAttribute(k_trsm<TRANS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  dim3 grid((NR + NB - 1) / NB, g.Kc);
  k_trsm<TRANS><<<grid, kSetupThreads, smem, st>>>(g.M, NR, Lc, src, dst, src_t);
  return cudaGetLastError();
}

}  // namespace

#define GPODE_CK(x)                      \
  do {                                   \
    cudaError_t e_ = (x);                \
    if (e_ != cudaSuccess) return e_;    \
  } while (0)

// (sizes in FLOATS of the caller's buffers: the factors and right-hand sides are stored in `real` = double)
size_t nu_save_floats(const NuGeom& g) { return (sizeof(real) / 4) * (static_cast<size_t>(g.Kc) * g.M * g.M + static_cast<size_t>(g.Kc) * g.M * g.NR); }
size_t nu_ws_floats(const NuGeom& g) {
  return (sizeof(real) / 4) * (2 * static_cast<size_t>(g.Kc) * g.M * g.M + 4 * static_cast<size_t>(g.Kc) * g.M * g.NR) + 64;
}

cudaError_t nu_forward(const NuGeom& g, const float* Z, const float* ell, const float* var, const float* u_prior, const float* u, float* nu,
                       float* save_f, int* info, float* ws_f, cudaStream_t st) {
  real* save = reinterpret_cast<real*>(save_f);
  real* Lc = save;                                                   // (Kc, M, M)
  real* a = save + static_cast<size_t>(g.Kc) * g.M * g.M;            // (Kc, M, NR): Lc^-1 u_prior
  real* b = reinterpret_cast<real*>(ws_f);                           // (Kc, M, NR)
  const long mm = static_cast<long>(g.M) * g.M, rhs = static_cast<long>(g.Kc) * g.M * g.NR;
  k_kzz_build<<<dim3(static_cast<unsigned>((mm + 255) / 256), g.Kc), 256, 0, st>>>(g, Z, ell, var, Lc);
  GPODE_CK(cudaGetLastError());
  const size_t cs = chol_smem(g.M);
  GPODE_CK(cudaFuncSetAttribute(k_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cs)));
  k_chol<<<g.Kc, kSetupThreads, cs, st>>>(g.M, Lc, info);
  GPODE_CK(cudaGetLastError());
  k_gather_lmd<<<blocks_for(rhs, 256), 256, 0, st>>>(g, u_prior, a);
  GPODE_CK(cudaGetLastError());
  GPODE_CK(trsm<false>(g, g.NR, Lc, a, a, 0, st));
  GPODE_CK(cudaMemcpyAsync(b, a, rhs * sizeof(real), cudaMemcpyDeviceToDevice, st));
  k_u_minus_a<<<blocks_for(rhs, 256), 256, 0, st>>>(g, u, b);
  GPODE_CK(cudaGetLastError());
  GPODE_CK(trsm<true>(g, g.NR, Lc, b, b, 0, st));
  k_scatter<<<blocks_for(rhs, 256), 256, 0, st>>>(g, b, nu, 0, 1.f);
  return cudaGetLastError();
}

cudaError_t nu_backward(const NuGeom& g, const float* Z, const float* ell, const float* var, const float* u, const float* save_f, const float* dnu,
                        float* d_uprior, float* d_u, float* d_Z, float* d_ell, float* d_var, float* ws_f, cudaStream_t st) {
  const real* save = reinterpret_cast<const real*>(save_f);
  const real* Lc = save;
  const real* a = save + static_cast<size_t>(g.Kc) * g.M * g.M;
  const size_t mm = static_cast<size_t>(g.Kc) * g.M * g.M, rhs = static_cast<size_t>(g.Kc) * g.M * g.NR;
  real* S = reinterpret_cast<real*>(ws_f);   // (Kc, M, M)
  real* Y = S + mm;              // (Kc, M, M)
  real* bb = Y + mm;             // (Kc, M, NR)
  real* rr = bb + rhs;
  real* qq = rr + rhs;
  k_gather_nu<<<blocks_for(static_cast<long>(rhs), 256), 256, 0, st>>>(g, dnu, bb);
  GPODE_CK(cudaGetLastError());
  GPODE_CK(trsm<false>(g, g.NR, Lc, bb, bb, 0, st));                 // bb = Lc^-1 nu_bar   (= u_bar)
  GPODE_CK(trsm<true>(g, g.NR, Lc, bb, rr, 0, st));                  // r  = Lc^-T bb       (= -u_prior_bar)
  GPODE_CK(trsm<true>(g, g.NR, Lc, a, qq, 0, st));                   // q  = Lc^-T a
  if (d_u) {
    k_scatter<<<blocks_for(static_cast<long>(rhs), 256), 256, 0, st>>>(g, bb, d_u, 1, 1.f);
    GPODE_CK(cudaGetLastError());
  }
  if (d_uprior) {
    k_scatter<<<blocks_for(static_cast<long>(rhs), 256), 256, 0, st>>>(g, rr, d_uprior, 1, -1.f);
    GPODE_CK(cudaGetLastError());
  }
  if (!d_Z && !d_ell && !d_var) return cudaSuccess;
  const long m2 = static_cast<long>(g.M) * g.M;
  k_build_S<<<dim3(static_cast<unsigned>((m2 + 255) / 256), g.Kc), 256, 0, st>>>(g, u, bb, S);
  GPODE_CK(cudaGetLastError());
  GPODE_CK(trsm<true>(g, g.M, Lc, S, Y, 0, st));                     // Y = Lc^-T S
  GPODE_CK(trsm<true>(g, g.M, Lc, Y, S, 1, st));                     // X^T = Lc^-T Y^T  (X symmetric) -> S
  if (d_Z) GPODE_CK(cudaMemsetAsync(d_Z, 0, static_cast<size_t>(g.M) * g.D_in * 4, st));
  if (d_ell) GPODE_CK(cudaMemsetAsync(d_ell, 0, static_cast<size_t>(g.dimwise ? g.D_out * g.D_in : g.D_in) * 4, st));
  if (d_var) GPODE_CK(cudaMemsetAsync(d_var, 0, static_cast<size_t>(g.dimwise ? g.D_out : 1) * 4, st));
  k_kzz_bwd<<<dim3((g.M + 7) / 8, g.Kc), 256, 0, st>>>(g, Z, ell, var, S, rr, qq, d_Z, d_ell, d_var);
  return cudaGetLastError();
}

cudaError_t inducing_forward(int L, int M, int D, const float* Lq, const float* Um, const float* eps, float* u, cudaStream_t st) {
  const long warps = static_cast<long>(M) * D;
  k_inducing_fwd<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, st>>>(L, M, D, Lq, Um, eps, u);
  return cudaGetLastError();
}
cudaError_t inducing_backward(int L, int M, int D, const float* eps, const float* du, float* dLq, float* dUm, int accumulate, cudaStream_t st) {
  const long n = static_cast<long>(M) * (M + 1) / 2 * D;
  k_inducing_bwd<<<blocks_for(n, 256, 148 * 16), 256, 0, st>>>(L, M, D, eps, du, dLq, dUm, accumulate);
  return cudaGetLastError();
}
cudaError_t kl_forward(int M, int D, const float* Lq, const float* Um, float* kl, cudaStream_t st) {
  GPODE_CK(cudaMemsetAsync(kl, 0, 4, st));
  const long n = static_cast<long>(M) * (M + 1) / 2 * D;
  k_kl_fwd<<<blocks_for(n, 256, 148 * 4), 256, 0, st>>>(M, D, Lq, Um, kl);
  return cudaGetLastError();
}
cudaError_t kl_backward(int M, int D, const float* Lq, const float* Um, const float* dkl, float* dLq, float* dUm, cudaStream_t st) {
  const long n = static_cast<long>(M) * (M + 1) / 2 * D;
  k_kl_bwd<<<blocks_for(n, 256, 148 * 8), 256, 0, st>>>(M, D, Lq, Um, dkl, dLq, dUm);
  GPODE_CK(cudaGetLastError());
  k_kl_bwd_diag<<<(M * D + 255) / 256, 256, 0, st>>>(M, D, Lq, dkl, dLq);
  return cudaGetLastError();
}

}  // namespace gpode
