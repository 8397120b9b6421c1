// Per-rollout setup of the sparse GP at the M inducing points, batched over output dimensions and MC samples
// (sm_100a): the pieces of SVGP_Layer.build_cache / compute_nu / kl that sit in the autograd graph
// (reference experiments/model/core/svpy.py:88-121,144-175; core/kernels.py:98-110,155-172).
//
//   nu = Lc^-T (u - Lc^-1 u_prior),  Lc = chol(K(Z,Z) + 1e-5 I)        one (M x M) system per output dim (RBF dimwise),
//                                                                       one shared by all output dims (RBF shared), or ONE
//                                                                       (M D x M D) system for the divergence-free kernel
//                                                                       (core/kernels.py:289-303,376-387), shared by the L samples
// Layout used inside: right-hand sides are (Kc, n, NR) row-major, Kc = number of matrices, n = their order (M or M D), NR =
// columns per matrix (dimwise / DF: the L samples; shared: L * D_out).  Matrices are (Kc, n, n) row-major; after the Cholesky the
// lower triangle holds Lc and the strict upper triangle is garbage (never read).  Like torch.linalg.cholesky the factorisation
// reads the LOWER triangle only (the DF Gram matrix is not symmetric once the (D,D) lengthscales / per-column variances differ,
// SURVEY.md Appendix B.5) and its backward hands the symmetrised gradient to BOTH triangles of K.
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
  if (g.df) return static_cast<size_t>(r) * g.n + i;            // (L, M, D) flattened per sample: row i = m * D + component (kernels.py:384-386)
  const int l = g.dimwise ? r : r / g.D_out;
  const int kk = g.dimwise ? k : r - l * g.D_out;
  return (static_cast<size_t>(l) * g.M + i) * g.D_out + kk;
}
// element (k, i, r) of nu in its reference layout: dimwise (L, D_out, M, 1); shared (L, M, D_out)
__device__ __forceinline__ size_t nu_index(const NuGeom& g, int k, int i, int r) {
  if (g.df) return static_cast<size_t>(r) * g.n + i;            // (L, M D, 1)
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
// blocked triangular solve with NC right-hand-side columns per CTA, RHS block resident in shared memory:
//   TRANS = false:  Lc X = B  (top to bottom)        TRANS = true:  Lc^T X = B  (bottom to top)
// src / dst are (Kc, M, NR) arrays (dst may alias src); src_t: read the source transposed ((Kc, NR, M), NR == M).
// NC = 32 for the small RBF systems; 16 / 8 / 4 keep the resident block within shared memory for the (M D)-order DF system
// (lane = (row in a group of 32 / NC, column)); `strip` rows of the off-diagonal panel are staged per pass (64, or 32 when tight).
// ---------------------------------------------------------------------------------------------
template <bool TRANS, int NC>
__global__ void __launch_bounds__(kSetupThreads) k_trsm(const int M, const int NR, const real* __restrict__ Lall, const real* src,
                                                        real* dst, const int src_t, const int strip) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  real* sm = reinterpret_cast<real*>(sm_raw);
  const int Mp = (M + NB - 1) / NB * NB;
  constexpr int XS = NC + 1, RPW = 32 / NC;
  real* Xs = sm;                          // [Mp][NC+1]
  real* Lb = Xs + Mp * XS;                // [NB][NB+1] diagonal block
  real* Ls = Lb + NB * (NB + 1);          // [strip][NB+1] rows of the off-diagonal strip
  const int k = blockIdx.y, c0 = blockIdx.x * NC;
  const int ncol = min(NC, NR - c0);
  const real* Lc = Lall + static_cast<size_t>(k) * M * M;
  const real* S = src + static_cast<size_t>(k) * M * NR;
  real* Dd = dst + static_cast<size_t>(k) * M * NR;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int col = lane % NC, rsub = lane / NC;
  if (src_t) {
    for (int e = tid; e < Mp * NC; e += blockDim.x) {
      const int c = e / Mp, i = e - c * Mp;                        // consecutive threads walk i: coalesced rows of the source
      Xs[i * XS + c] = (i < M && c < ncol) ? S[static_cast<size_t>(c0 + c) * M + i] : real(0);
    }
  } else {
    for (int e = tid; e < Mp * NC; e += blockDim.x) {
      const int i = e / NC, c = e - i * NC;
      Xs[i * XS + c] = (i < M && c < ncol) ? S[static_cast<size_t>(i) * NR + c0 + c] : real(0);
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
    if (warp == 0 && lane < NC) {          // lane = column: 32-step substitution on the diagonal block
      real x[NB];
#pragma unroll
      for (int r = 0; r < NB; ++r) x[r] = Xs[(j0 + r) * XS + lane];
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
      for (int r = 0; r < NB; ++r) Xs[(j0 + r) * XS + lane] = x[r];
    }
    __syncthreads();
    // update the remaining rows: non-trans rows i >= j0 + NB use Lc[i][j0 + c]; trans rows j < j0 use Lc[j0 + c][j]
    const int lo_row = TRANS ? 0 : j0 + NB, hi_row = TRANS ? j0 : Mp;
    for (int s0 = lo_row; s0 < hi_row; s0 += strip) {
      const int ns = min(strip, hi_row - s0);
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
      for (int rr = warp * RPW + rsub; rr < ns; rr += nwarp * RPW) {
        real acc = 0;
#pragma unroll
        for (int c = 0; c < NB; ++c) acc = fma(Ls[rr * (NB + 1) + c], Xs[(j0 + c) * XS + col], acc);
        Xs[(s0 + rr) * XS + col] -= acc;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  for (int e = tid; e < M * NC; e += blockDim.x) {
    const int i = e / NC, c = e - i * NC;
    if (c < ncol) Dd[static_cast<size_t>(i) * NR + c0 + c] = Xs[i * XS + c];
  }
}

// ---------------------------------------------------------------------------------------------
// One step of a right-looking blocked Cholesky of ONE large matrix (the (M D)-order DF system) spread over the chip:
// launch j0 = 0, NB, 2 NB ... (stream order is the only synchronisation).  Every CTA factors the current NB x NB diagonal
// block itself (redundantly: 32 dependent column steps, cheaper than a grid-wide barrier), turns the panel rows of ITS two
// 64-row tiles into columns of Lc (P = A[rows, j0..] D^-T) and subtracts P_I P_K^T from its tile of the trailing lower
// triangle.  A is the working matrix (only its trailing part is ever written), Lc a separate output -- so no CTA reads
// what another one writes within a launch.  CTAs of the first tile column store their panel rows, CTA 0 the diagonal block.
// ---------------------------------------------------------------------------------------------
constexpr int CT = 64;
__global__ void __launch_bounds__(kSetupThreads) k_chol_step(const int n, const int j0, real* __restrict__ A, real* __restrict__ Lc, int* __restrict__ info) {
  __shared__ real Dg[NB * (NB + 1)];
  __shared__ real Rd[NB];
  __shared__ real Pt[2][NB][CT + 2];     // panel rows of the I tile / K tile, transposed
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbk = min(NB, n - j0);
  int ti = static_cast<int>((sqrtf(8.f * static_cast<float>(blockIdx.x) + 1.f) - 1.f) * 0.5f);
  while (ti * (ti + 1) / 2 > static_cast<int>(blockIdx.x)) --ti;
  while ((ti + 1) * (ti + 2) / 2 <= static_cast<int>(blockIdx.x)) ++ti;
  const int tk = static_cast<int>(blockIdx.x) - ti * (ti + 1) / 2;
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e - r * NB;
    Dg[r * (NB + 1) + c] = (r < nbk && c < nbk && c <= r) ? A[static_cast<size_t>(j0 + r) * n + j0 + c] : (r == c ? real(1) : real(0));
  }
  __syncthreads();
  {   // every warp factors the block (lane = row, held in registers; column c is broadcast by shuffles): no divergent branch around
      // the shuffles, no shared-memory round trips; warp 0 stores the result
    int bad = 0;
    real x[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) x[c] = Dg[lane * (NB + 1) + c];
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      const real piv = __shfl_sync(0xffffffffu, x[c], c);
      if (!(piv > real(0)) && c < nbk && bad == 0) bad = j0 + c + 1;
      const real rs = rsqrt(piv);
      x[c] = lane == c ? piv * rs : x[c] * rs;
      if (lane == c && warp == 0) Rd[c] = rs;                        // 1 / Lc[c][c] for the panel rows
#pragma unroll
      for (int cc = c + 1; cc < NB; ++cc) {
        const real lcc = __shfl_sync(0xffffffffu, x[c], cc);
        x[cc] = lane >= cc ? fma(-x[c], lcc, x[cc]) : x[cc];
      }
    }
    if (warp == 0) {
#pragma unroll
      for (int c = 0; c < NB; ++c) Dg[lane * (NB + 1) + c] = c <= lane ? x[c] : real(0);
      if (blockIdx.x == 0 && lane == 0 && info && bad != 0 && *info == 0) *info = bad;
    }
  }
  __syncthreads();
  if (blockIdx.x == 0)
    for (int e = tid; e < nbk * nbk; e += blockDim.x) {
      const int r = e / nbk, c = e - r * nbk;
      if (c <= r) Lc[static_cast<size_t>(j0 + r) * n + j0 + c] = Dg[r * (NB + 1) + c];
    }
  const int r0 = j0 + NB;
  if (r0 >= n) return;
  const int rowI = r0 + ti * CT, rowK = r0 + tk * CT;
  if (tid < 2 * CT) {      // one thread per panel row: forward substitution over the NB columns
    const int which = tid / CT, rr = tid - which * CT;
    const int grow = (which ? rowK : rowI) + rr;
    real x[NB];
    if (grow < n) {
      const real* arow = A + static_cast<size_t>(grow) * n + j0;
#pragma unroll
      for (int c = 0; c < NB; ++c) x[c] = c < nbk ? arow[c] : real(0);
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        real v = x[c];
#pragma unroll
        for (int cc = 0; cc < c; ++cc) v = fma(-x[cc], Dg[c * (NB + 1) + cc], v);
        x[c] = v * Rd[c];
      }
      if (which == 0 && tk == 0) {
        real* lrow = Lc + static_cast<size_t>(grow) * n + j0;
#pragma unroll
        for (int c = 0; c < NB; ++c)
          if (c < nbk) lrow[c] = x[c];
      }
    } else {
#pragma unroll
      for (int c = 0; c < NB; ++c) x[c] = real(0);
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) Pt[which][c][rr] = x[c];
  }
  __syncthreads();
  // trailing update of this tile: 16 x 16 threads, 4 x 4 entries each
  const int ty = tid >> 4, tx = tid & 15;
  real acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = real(0);
#pragma unroll 4
  for (int c = 0; c < NB; ++c) {
    real vi[4], vj[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      vi[a] = Pt[0][c][4 * ty + a];
      vj[a] = Pt[1][c][4 * tx + a];
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fma(vi[a], vj[b], acc[a][b]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = rowI + 4 * ty + a;
    if (i >= n) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = rowK + 4 * tx + b;
      if (j <= i) A[static_cast<size_t>(i) * n + j] -= acc[a][b];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Explicit inverse of the (large, single) Cholesky factor: Linv = Lc^-1, lower triangular.  Every later solve of the DF setup
// -- two skinny ones in the forward, three skinny and two (n x n) ones in the backward -- then is a plain matrix product with
// no sequential dependency (substitution on ONE matrix is a chain of n / 32 dependent steps whatever the number of SMs).
//   level 0: the NB x NB diagonal blocks are inverted (one warp each: lane = column of the inverse);
//   level l: pairs of finished diagonal blocks [A 0; B C] of size b = NB 2^l are joined: Linv[B part] = - C^-1 (B A^-1),
//            two batched products per level (k_dgemm), log2(n / NB) levels.
// fp64 throughout: cond(Lc) = sqrt(cond(K)) <= ~1e3 at the reference's settings, the products stay at 1e-12.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_diag_inv(const int n, const real* __restrict__ Lc, real* __restrict__ Linv) {
  __shared__ real Dg[NB * (NB + 1)];
  const int j0 = blockIdx.x * NB, lane = threadIdx.x;
  const int nbk = min(NB, n - j0);
  for (int e = lane; e < NB * NB; e += 32) {
    const int r = e / NB, c = e - r * NB;
    Dg[r * (NB + 1) + c] = (r < nbk && c < nbk && c <= r) ? Lc[static_cast<size_t>(j0 + r) * n + j0 + c] : (r == c ? real(1) : real(0));
  }
  __syncwarp();
  real x[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    real v0 = r == lane ? real(1) : real(0), v1 = 0;
#pragma unroll
    for (int c = 0; c < r; ++c) {
      if (c & 1) v1 = fma(-Dg[r * (NB + 1) + c], x[c], v1);
      else v0 = fma(-Dg[r * (NB + 1) + c], x[c], v0);
    }
    x[r] = (v0 + v1) / Dg[r * (NB + 1) + r];
  }
  if (lane < nbk) {
#pragma unroll
    for (int r = 0; r < NB; ++r)
      if (r < nbk) Linv[static_cast<size_t>(j0 + r) * n + j0 + lane] = r >= lane ? x[r] : real(0);
  }
}

// out (n, NR) = Linv x (TRANS = 0) or Linv^T x (TRANS = 1) for a lower-triangular Linv (n, n) and NR <= 8 columns: the skinny solves of
// the DF setup as matrix-vector products.  TRANS = 0: one warp per output row (lanes stride the row, k <= i); TRANS = 1: a CTA owns
// 32 output columns j, its 8 warps stride the rows i >= j of Linv (coalesced along j), partial sums meet in shared memory.
template <int TRANS>
__global__ void __launch_bounds__(256) k_tri_matvec(const int n, const int NR, const real* __restrict__ Linv, const real* __restrict__ x, real* __restrict__ out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  real acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0;
  if (!TRANS) {
    const int i = blockIdx.x * 8 + warp;
    if (i >= n) return;
    const real* row = Linv + static_cast<size_t>(i) * n;
    for (int k = lane; k <= i; k += 32) {
      const real a = row[k];
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r < NR) acc[r] = fma(a, x[static_cast<size_t>(k) * NR + r], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < NR) {
        const real v = warp_sum_real(acc[r]);
        if (lane == 0) out[static_cast<size_t>(i) * NR + r] = v;
      }
  } else {
    __shared__ real part[8][8][33];
    const int j = blockIdx.x * 32 + lane;
    if (j < n)
      for (int i = blockIdx.x * 32 + warp; i < n; i += 8) {
        if (i < j) continue;
        const real a = Linv[static_cast<size_t>(i) * n + j];
#pragma unroll
        for (int r = 0; r < 8; ++r)
          if (r < NR) acc[r] = fma(a, x[static_cast<size_t>(i) * NR + r], acc[r]);
      }
#pragma unroll
    for (int r = 0; r < 8; ++r) part[warp][r][lane] = acc[r];
    __syncthreads();
    if (warp < NR && j < n) {
      real v = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += part[w][warp][lane];
      out[static_cast<size_t>(j) * NR + warp] = v;
    }
  }
}

// C = alpha op(A) B (+ beta C), row-major fp64, batched over blockIdx.z with element strides; op(A) = A^T when ta (A stored K x M).
// Batch p uses M_eff = min(M, m_lim - p * m_lim_step) rows (ragged last block of the level-wise inverse; K_eff = M_eff when k_is_m).
// 64 x 64 tile, 256 threads, 4 x 4 per thread, K in chunks of 16.
constexpr int GT = 64, GK = 16;
__global__ void __launch_bounds__(256) k_dgemm(int M, const int N, int K, const real alpha, const real* __restrict__ A, const int lda, const long sa, const int ta,
                                               const real* __restrict__ B, const int ldb, const long sb, const real beta, real* __restrict__ C, const int ldc,
                                               const long sc, const int m_lim, const int m_lim_step, const int k_is_m, const int tri) {
  // tri (operand structure, skips all-zero K chunks): 1 = B lower triangular (k >= j), 2 = op(A) = A^T with A lower triangular (k >= i),
  //                                                   4 = A lower triangular, not transposed (k <= i)
  __shared__ real As[GK][GT + 4], Bs[GK][GT + 4];
  const int p = blockIdx.z;
  if (m_lim_step > 0 || m_lim > 0) {
    const int lim = m_lim - p * m_lim_step;
    if (lim < M) M = lim;
    if (M <= 0) return;
    if (k_is_m) K = M;
  }
  A += p * sa;
  B += p * sb;
  C += p * sc;
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  if (i0 >= M || j0 >= N) return;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  real acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0;
  int k_begin = 0, k_end = K;
  if (tri & 1) k_begin = max(k_begin, j0 / GK * GK);
  if (tri & 2) k_begin = max(k_begin, i0 / GK * GK);
  if (tri & 4) k_end = min(k_end, i0 + GT);
  // software pipeline: the global loads of chunk k0 + GK are in flight (registers) while chunk k0 is multiplied out of shared memory
  real ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = tid + 256 * t;
      int kk, ii;
      if (ta) { kk = e / GT; ii = e - kk * GT; } else { ii = e / GK; kk = e - ii * GK; }     // consecutive threads walk the contiguous direction
      const int gi = i0 + ii, gk = k0 + kk;
      ra[t] = (gi < M && gk < K) ? (ta ? A[static_cast<size_t>(gk) * lda + gi] : A[static_cast<size_t>(gi) * lda + gk]) : real(0);
      const int kb = e / GT, jj = e - kb * GT;
      const int gkb = k0 + kb, gj = j0 + jj;
      rb[t] = (gkb < K && gj < N) ? B[static_cast<size_t>(gkb) * ldb + gj] : real(0);
    }
  };
  if (k_begin < k_end) fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += GK) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = tid + 256 * t;
      int kk, ii;
      if (ta) { kk = e / GT; ii = e - kk * GT; } else { ii = e / GK; kk = e - ii * GK; }
      As[kk][ii] = ra[t];
      Bs[e / GT][e % GT] = rb[t];
    }
    __syncthreads();
    if (k0 + GK < k_end) fetch(k0 + GK);
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      real va[4], vb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        va[a] = As[kk][4 * ty + a];
        vb[a] = Bs[kk][4 * tx + a];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(va[a], vb[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gi = i0 + 4 * ty + a;
    if (gi >= M) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gj = j0 + 4 * tx + b;
      if (gj >= N) continue;
      real* c = C + static_cast<size_t>(gi) * ldc + gj;
      *c = beta == real(0) ? alpha * acc[a][b] : fma(alpha, acc[a][b], beta * *c);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// divergence-free Gram matrix K(Z,Z) + jitter I (core/kernels.py:289-303): row m D + i, column m' D + j,
//   d = Z_m' - Z_m, c_ij = 1 / ell_ij^2:   var_j c_ij exp(-r2 c_ij / 2) (d_i d_j c_ij + delta_ij ((D - 1) - r2 c_ij))
// one thread per (m, m') block.
// ---------------------------------------------------------------------------------------------
__global__ void k_df_kzz_build(const NuGeom g, const float* __restrict__ Z, const float* __restrict__ ell, const float* __restrict__ var,
                               real* __restrict__ A) {
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long>(g.M) * g.M) return;
  const int D = g.D_in;
  const int m = static_cast<int>(idx / g.M), mp = static_cast<int>(idx - static_cast<long>(m) * g.M);
  real d[8], r2 = 0;
  for (int k = 0; k < D; ++k) {
    d[k] = static_cast<real>(Z[mp * D + k]) - static_cast<real>(Z[m * D + k]);
    r2 = fma(d[k], d[k], r2);
  }
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      const real e = ell[i * D + j], c = 1.0 / (e * e);
      real v = static_cast<real>(var[j]) * c * exp(-0.5 * r2 * c) * (d[i] * d[j] * c + (i == j ? (D - 1.0) - r2 * c : 0.0));
      if (m == mp && i == j) v += static_cast<real>(g.jitter);
      A[(static_cast<size_t>(m) * D + i) * g.n + static_cast<size_t>(mp) * D + j] = v;
    }
}
// A_bar = X + 1/2 sum_r (r_i q_j + q_i r_j) (symmetric; torch hands it to both triangles of K) contracted with dK/d(var, ell, Z)
// of the divergence-free Gram matrix.  One CTA per inducing point m, threads over m': each thread contracts the (m, m') block
// for the lengthscale / variance sums and the Z_m share of it (d = Z_m' - Z_m: -g), plus the (m', m) block for Z_m's share as
// the COLUMN point (+g) -- so dZ needs no atomics.  acc (doubles): [D*D] lengthscales, then [D] variances.
__global__ void __launch_bounds__(128) k_df_kzz_bwd(const NuGeom g, const float* __restrict__ Z, const float* __restrict__ ell, const float* __restrict__ var,
                                                     const real* __restrict__ X, const real* __restrict__ rr_, const real* __restrict__ qq_,
                                                     float* __restrict__ d_Z, real* __restrict__ acc) {
  const int D = g.D_in, m = blockIdx.x, tid = threadIdx.x;
  __shared__ real red[128];
  real c_[64], dl[64], dv[8], dz[8];
  for (int e = 0; e < D * D; ++e) {
    const real l = ell[e];
    c_[e] = 1.0 / (l * l);
    dl[e] = 0;
  }
  for (int k = 0; k < D; ++k) dv[k] = dz[k] = 0;
  for (int mp = tid; mp < g.M; mp += blockDim.x) {
    real d[8], r2 = 0;
    for (int k = 0; k < D; ++k) {
      d[k] = static_cast<real>(Z[mp * D + k]) - static_cast<real>(Z[m * D + k]);
      r2 = fma(d[k], d[k], r2);
    }
    for (int pass = 0; pass < 2; ++pass) {
      // pass 0: block (m, m') with d;   pass 1: block (m', m) with -d (only its Z_m share is taken)
      const int ra = pass == 0 ? m : mp, ca = pass == 0 ? mp : m;
      const real sgn = pass == 0 ? 1.0 : -1.0;
      for (int i = 0; i < D; ++i) {
        const size_t row = static_cast<size_t>(ra) * D + i;
        for (int j = 0; j < D; ++j) {
          const size_t colx = static_cast<size_t>(ca) * D + j;
          real ab = X[row * g.n + colx];
          real rk = 0;
          for (int r = 0; r < g.NR; ++r) rk += rr_[row * g.NR + r] * qq_[colx * g.NR + r] + qq_[row * g.NR + r] * rr_[colx * g.NR + r];
          ab = fma(0.5, rk, ab);
          const real c = c_[i * D + j], E = exp(-0.5 * r2 * c), vj = var[j];
          const real di = sgn * d[i], dj = sgn * d[j];
          const real H = di * dj * c + (i == j ? (D - 1.0) - r2 * c : 0.0);
          if (pass == 0) {
            dv[j] = fma(ab, c * E * H, dv[j]);
            // d/dc: var E [H + c (-r2/2 H + d_i d_j - delta_ij r2)]
            dl[i * D + j] = fma(ab, vj * E * (H + c * (-0.5 * r2 * H + di * dj - (i == j ? r2 : 0.0))), dl[i * D + j]);
          }
          // d/dd_k of var c E H (d = column point - row point): var c E [ -c d_k H + c (delta_ik d_j + delta_jk d_i) - 2 delta_ij c d_k ]
          const real base = ab * vj * c * E * c;
          for (int k = 0; k < D; ++k) {
            const real dk = sgn * d[k];
            real t = -dk * H - (i == j ? 2.0 * dk : 0.0);
            if (k == i) t += dj;
            if (k == j) t += di;
            // Z_m is the row point in pass 0 (d/dZ_m = -d/dd) and the column point in pass 1 (+d/dd)
            dz[k] = fma(pass == 0 ? -base : base, t, dz[k]);
          }
        }
      }
    }
  }
  auto block_sum = [&](real v) -> real {
    v = warp_sum_real(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    real s = 0;
    if (tid == 0)
      for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) s += red[w];
    return s;
  };
  for (int k = 0; k < D; ++k) {
    const real s = block_sum(dz[k]);
    if (tid == 0 && d_Z) d_Z[m * D + k] = static_cast<float>(s);
  }
  for (int e = 0; e < D * D; ++e) {
    const real s = block_sum(dl[e]);
    if (tid == 0) atomicAdd(&acc[e], s * (-2.0 * c_[e] / static_cast<real>(ell[e])));     // dc/dell = -2 c / ell
  }
  for (int k = 0; k < D; ++k) {
    const real s = block_sum(dv[k]);
    if (tid == 0) atomicAdd(&acc[D * D + k], s);
  }
}
__global__ void k_df_acc_out(const int D, const real* __restrict__ acc, float* __restrict__ d_ell, float* __restrict__ d_var) {
  const int e = threadIdx.x;
  if (e < D * D && d_ell) d_ell[e] = static_cast<float>(acc[e]);
  if (e < D && d_var) d_var[e] = static_cast<float>(acc[D * D + e]);
}

// ---------------------------------------------------------------------------------------------
// small glue kernels (all elementwise over (k, i, r) or (k, i, j))
// ---------------------------------------------------------------------------------------------
// rhs[k][i][r] = reference-layout (L, M, D_out) tensor
__global__ void k_gather_lmd(const NuGeom g, const float* __restrict__ src, real* __restrict__ rhs) {
  const long n = static_cast<long>(g.Kc) * g.n * g.NR;
  for (long e = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += static_cast<long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(e % g.NR), i = static_cast<int>((e / g.NR) % g.n), k = static_cast<int>(e / (static_cast<long>(g.NR) * g.n));
    rhs[e] = src[lmd_index(g, k, i, r)];
  }
}
// b = u - a  (u in reference layout), in place on the a array
__global__ void k_u_minus_a(const NuGeom g, const float* __restrict__ u, real* __restrict__ a_then_b) {
  const long n = static_cast<long>(g.Kc) * g.n * g.NR;
  for (long e = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += static_cast<long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(e % g.NR), i = static_cast<int>((e / g.NR) % g.n), k = static_cast<int>(e / (static_cast<long>(g.NR) * g.n));
    a_then_b[e] = static_cast<real>(u[lmd_index(g, k, i, r)]) - a_then_b[e];
  }
}
// scatter (Kc, M, NR) -> nu layout (mode 0) or (L, M, D_out) layout with a sign (mode 1)
__global__ void k_scatter(const NuGeom g, const real* __restrict__ rhs, float* __restrict__ out, const int mode, const float scale) {
  const long n = static_cast<long>(g.Kc) * g.n * g.NR;
  for (long e = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += static_cast<long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(e % g.NR), i = static_cast<int>((e / g.NR) % g.n), k = static_cast<int>(e / (static_cast<long>(g.NR) * g.n));
    out[mode == 0 ? nu_index(g, k, i, r) : lmd_index(g, k, i, r)] = static_cast<float>(scale * rhs[e]);
  }
}
// gather nu-layout tensor into (Kc, M, NR)
__global__ void k_gather_nu(const NuGeom g, const float* __restrict__ src, real* __restrict__ rhs) {
  const long n = static_cast<long>(g.Kc) * g.n * g.NR;
  for (long e = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += static_cast<long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(e % g.NR), i = static_cast<int>((e / g.NR) % g.n), k = static_cast<int>(e / (static_cast<long>(g.NR) * g.n));
    rhs[e] = src[nu_index(g, k, i, r)];
  }
}
// S[k][i][j] = -1/2 sum_r u[k][max(i,j)][r] bb[k][min(i,j)][r]
__global__ void k_build_S(const NuGeom g, const float* __restrict__ u, const real* __restrict__ bb, real* __restrict__ S) {
  const int k = blockIdx.y;
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long>(g.n) * g.n) return;
  const int i = static_cast<int>(idx / g.n), j = static_cast<int>(idx - static_cast<long>(i) * g.n);
  const int hi_ = i > j ? i : j, lo_ = i > j ? j : i;
  real acc = 0;
  for (int r = 0; r < g.NR; ++r) acc = fma(static_cast<real>(u[lmd_index(g, k, hi_, r)]), bb[(static_cast<size_t>(k) * g.n + lo_) * g.NR + r], acc);
  S[(static_cast<size_t>(k) * g.n + i) * g.n + j] = -0.5 * acc;
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
inline size_t trsm_smem(int M, int nc, int strip) {
  const int Mp = (M + NB - 1) / NB * NB;
  return (static_cast<size_t>(Mp) * (nc + 1) + NB * (NB + 1) + static_cast<size_t>(strip) * (NB + 1)) * sizeof(real);
}
inline size_t chol_smem(int M) {
  const int Mp = (M + NB - 1) / NB * NB;
  return (static_cast<size_t>(NB) * (NB + 1) + static_cast<size_t>(NB) * Mp) * sizeof(real);
}
constexpr size_t kSmemCap = 227 * 1024;

template <bool TRANS, int NC>
cudaError_t trsm_nc(int n, int Kc, int NR, int strip, const real* Lc, const real* src, real* dst, int src_t, cudaStream_t st) {
  const size_t smem = trsm_smem(n, NC, strip);
  cudaError_t e = cudaFuncSetAttribute(k_trsm<TRANS, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  dim3 grid((NR + NC - 1) / NC, Kc);
  k_trsm<TRANS, NC><<<grid, kSetupThreads, smem, st>>>(n, NR, Lc, src, dst, src_t, strip);
  return cudaGetLastError();
}
// widest column block whose resident right-hand sides fit in shared memory (and no wider than the columns there are)
template <bool TRANS>
cudaError_t trsm(const NuGeom& g, int NR, const real* Lc, const real* src, real* dst, int src_t, cudaStream_t st) {
  for (int nc = NR > 16 ? 32 : (NR > 8 ? 16 : 8); nc >= 4; nc >>= 1) {   // (few columns: a narrower block wastes no lanes)
    for (int strip = 64; strip >= 32; strip >>= 1) {
      if (trsm_smem(g.n, nc, strip) > kSmemCap) continue;
      switch (nc) {
        case 32: return trsm_nc<TRANS, 32>(g.n, g.Kc, NR, strip, Lc, src, dst, src_t, st);
        case 16: return trsm_nc<TRANS, 16>(g.n, g.Kc, NR, strip, Lc, src, dst, src_t, st);
        case 8: return trsm_nc<TRANS, 8>(g.n, g.Kc, NR, strip, Lc, src, dst, src_t, st);
        default: return trsm_nc<TRANS, 4>(g.n, g.Kc, NR, strip, Lc, src, dst, src_t, st);
      }
    }
  }
  return cudaErrorInvalidValue;
}
inline cudaError_t dgemm(int M, int N, int K, real alpha, const real* A, int lda, int ta, const real* B, int ldb, real beta, real* C, int ldc, cudaStream_t st,
                         int tri = 0, int batch = 1, long sa = 0, long sb = 0, long sc = 0, int m_lim = 0, int m_lim_step = 0, int k_is_m = 0) {
  dim3 grid((N + GT - 1) / GT, (M + GT - 1) / GT, batch);
  k_dgemm<<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, sa, ta, B, ldb, sb, beta, C, ldc, sc, m_lim, m_lim_step, k_is_m, tri);
  return cudaGetLastError();
}
inline cudaError_t tri_matvec(int n, int NR, const real* Linv, int trans, const real* x, real* out, cudaStream_t st) {
  if (NR > 8) return dgemm(n, NR, n, 1.0, Linv, n, trans, x, NR, 0.0, out, NR, st, trans ? 2 : 4);
  if (trans) k_tri_matvec<1><<<(n + 31) / 32, 256, 0, st>>>(n, NR, Linv, x, out);
  else k_tri_matvec<0><<<(n + 7) / 8, 256, 0, st>>>(n, NR, Linv, x, out);
  return cudaGetLastError();
}
// Linv = Lc^-1 (see k_diag_inv); T: (n, n) scratch
cudaError_t invert_factor(int n, const real* Lc, real* Linv, real* T, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(Linv, 0, static_cast<size_t>(n) * n * sizeof(real), st);
  if (e != cudaSuccess) return e;
  k_diag_inv<<<(n + NB - 1) / NB, 32, 0, st>>>(n, Lc, Linv);
  for (int b = NB; b < n; b <<= 1) {
    const int pairs = (n + 2 * b - 1) / (2 * b);                       // the last pair may have a short (or no) second block
    const long st_diag = 2L * b * (n + 1);
    // T[B part] = B A^-1      (rows r0 + b .. of Lc, columns r0 .. r0 + b)
    if ((e = dgemm(b, b, b, 1.0, Lc + static_cast<size_t>(b) * n, n, 0, Linv, n, 0.0, T + static_cast<size_t>(b) * n, n, st, 1, pairs, st_diag, st_diag, st_diag, n - b,
                   2 * b, 0)) != cudaSuccess)
      return e;
    // Linv[B part] = - C^-1 T
    if ((e = dgemm(b, b, b, -1.0, Linv + static_cast<size_t>(b) * (n + 1), n, 0, T + static_cast<size_t>(b) * n, n, 0.0, Linv + static_cast<size_t>(b) * n, n, st, 4, pairs,
                   st_diag, st_diag, st_diag, n - b, 2 * b, 1)) != cudaSuccess)
      return e;
  }
  return cudaGetLastError();
}

// Cholesky of the Kc matrices in `Lc` (in place for the batched small systems; through the working copy `work` for one large system)
cudaError_t cholesky(const NuGeom& g, real* Lc, real* work, int* info, cudaStream_t st) {
  const size_t cs = chol_smem(g.n);
  if (!g.df) {
    if (cs > kSmemCap) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cs));
    if (e != cudaSuccess) return e;
    k_chol<<<g.Kc, kSetupThreads, cs, st>>>(g.n, Lc, info);
    return cudaGetLastError();
  }
  // one large system: `work` holds K + jitter I, the factor goes to Lc
  if (info) {
    cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
  }
  for (int j0 = 0; j0 < g.n; j0 += NB) {
    const int rem = g.n - (j0 + NB);
    const int nt = rem > 0 ? (rem + CT - 1) / CT : 0;
    const int grid = nt > 0 ? nt * (nt + 1) / 2 : 1;
    k_chol_step<<<grid, kSetupThreads, 0, st>>>(g.n, j0, work, Lc, info);
  }
  return cudaGetLastError();
}

}  // namespace

#define GPODE_CK(x)                      \
  do {                                   \
    cudaError_t e_ = (x);                \
    if (e_ != cudaSuccess) return e_;    \
  } while (0)

// (sizes in FLOATS of the caller's buffers: the factors and right-hand sides are stored in `real` = double)
size_t nu_save_floats(const NuGeom& g) { return (sizeof(real) / 4) * (static_cast<size_t>(g.Kc) * g.n * g.n + static_cast<size_t>(g.Kc) * g.n * g.NR); }
size_t nu_ws_floats(const NuGeom& g) {
  return (sizeof(real) / 4) * ((g.df ? 3 : 2) * static_cast<size_t>(g.Kc) * g.n * g.n + 4 * static_cast<size_t>(g.Kc) * g.n * g.NR + 128) + 64;
}

// DF: one large system.  save = [ Linv (n, n) | a = Lc^-1 u_prior (n, NR) ]
static cudaError_t df_nu_forward(const NuGeom& g, const float* Z, const float* ell, const float* var, const float* u_prior, const float* u, float* nu,
                                 real* save, int* info, real* ws, cudaStream_t st) {
  const size_t nn = static_cast<size_t>(g.n) * g.n, rhs = static_cast<size_t>(g.n) * g.NR;
  real* Linv = save;
  real* a = save + nn;
  real* b = ws;              // (n, NR)
  real* pin = b + rhs;       // (n, NR) gathered u_prior
  real* work = pin + rhs;    // (n, n)  K + jitter I, consumed by the factorisation; then scratch of the inversion
  real* Lc = work + nn;      // (n, n)
  const long mm = static_cast<long>(g.M) * g.M;
  k_df_kzz_build<<<static_cast<unsigned>((mm + 127) / 128), 128, 0, st>>>(g, Z, ell, var, work);
  GPODE_CK(cudaGetLastError());
  GPODE_CK(cholesky(g, Lc, work, info, st));
  GPODE_CK(invert_factor(g.n, Lc, Linv, work, st));
  k_gather_lmd<<<blocks_for(static_cast<long>(rhs), 256), 256, 0, st>>>(g, u_prior, pin);
  GPODE_CK(cudaGetLastError());
  GPODE_CK(tri_matvec(g.n, g.NR, Linv, 0, pin, a, st));                                     // a = Lc^-1 u_prior
  GPODE_CK(cudaMemcpyAsync(b, a, rhs * sizeof(real), cudaMemcpyDeviceToDevice, st));
  k_u_minus_a<<<blocks_for(static_cast<long>(rhs), 256), 256, 0, st>>>(g, u, b);
  GPODE_CK(cudaGetLastError());
  GPODE_CK(tri_matvec(g.n, g.NR, Linv, 1, b, pin, st));                                     // nu = Lc^-T (u - a)
  k_scatter<<<blocks_for(static_cast<long>(rhs), 256), 256, 0, st>>>(g, pin, nu, 0, 1.f);
  return cudaGetLastError();
}

static cudaError_t df_nu_backward(const NuGeom& g, const float* Z, const float* ell, const float* var, const float* u, const real* save, const float* dnu,
                                  float* d_uprior, float* d_u, float* d_Z, float* d_ell, float* d_var, real* ws, cudaStream_t st) {
  const size_t nn = static_cast<size_t>(g.n) * g.n, rhs = static_cast<size_t>(g.n) * g.NR;
  const real* Linv = save;
  const real* a = save + nn;
  real* S = ws;
  real* Y = S + nn;
  real* bb = Y + nn;
  real* rr = bb + rhs;
  real* qq = rr + rhs;
  real* gin = qq + rhs;
  real* dacc = gin + rhs;
  k_gather_nu<<<blocks_for(static_cast<long>(rhs), 256), 256, 0, st>>>(g, dnu, gin);
  GPODE_CK(cudaGetLastError());
  GPODE_CK(tri_matvec(g.n, g.NR, Linv, 0, gin, bb, st));                                    // bb = Lc^-1 nu_bar   (= u_bar)
  GPODE_CK(tri_matvec(g.n, g.NR, Linv, 1, bb, rr, st));                                     // r  = Lc^-T bb       (= -u_prior_bar)
  GPODE_CK(tri_matvec(g.n, g.NR, Linv, 1, a, qq, st));                                      // q  = Lc^-T a
  if (d_u) {
    k_scatter<<<blocks_for(static_cast<long>(rhs), 256), 256, 0, st>>>(g, bb, d_u, 1, 1.f);
    GPODE_CK(cudaGetLastError());
  }
  if (d_uprior) {
    k_scatter<<<blocks_for(static_cast<long>(rhs), 256), 256, 0, st>>>(g, rr, d_uprior, 1, -1.f);
    GPODE_CK(cudaGetLastError());
  }
  if (!d_Z && !d_ell && !d_var) return cudaSuccess;
  const long m2 = static_cast<long>(g.n) * g.n;
  k_build_S<<<dim3(static_cast<unsigned>((m2 + 255) / 256), 1), 256, 0, st>>>(g, u, bb, S);
  GPODE_CK(cudaGetLastError());
  GPODE_CK(dgemm(g.n, g.n, g.n, 1.0, S, g.n, 0, Linv, g.n, 0.0, Y, g.n, st, 1));            // Y = S Lc^-1
  GPODE_CK(dgemm(g.n, g.n, g.n, 1.0, Linv, g.n, 1, Y, g.n, 0.0, S, g.n, st, 2));            // X = Lc^-T S Lc^-1 -> S
  GPODE_CK(cudaMemsetAsync(dacc, 0, 128 * sizeof(real), st));
  k_df_kzz_bwd<<<g.M, 128, 0, st>>>(g, Z, ell, var, S, rr, qq, d_Z, dacc);
  GPODE_CK(cudaGetLastError());
  k_df_acc_out<<<1, 64, 0, st>>>(g.D_in, dacc, d_ell, d_var);
  return cudaGetLastError();
}

cudaError_t nu_forward(const NuGeom& g, const float* Z, const float* ell, const float* var, const float* u_prior, const float* u, float* nu,
                       float* save_f, int* info, float* ws_f, cudaStream_t st) {
  real* save = reinterpret_cast<real*>(save_f);
  if (g.df) return df_nu_forward(g, Z, ell, var, u_prior, u, nu, save, info, reinterpret_cast<real*>(ws_f), st);
  real* Lc = save;                                                   // (Kc, M, M)
  real* a = save + static_cast<size_t>(g.Kc) * g.n * g.n;            // (Kc, M, NR): Lc^-1 u_prior
  real* b = reinterpret_cast<real*>(ws_f);                           // (Kc, M, NR)
  const long mm = static_cast<long>(g.M) * g.M, rhs = static_cast<long>(g.Kc) * g.n * g.NR;
  k_kzz_build<<<dim3(static_cast<unsigned>((mm + 255) / 256), g.Kc), 256, 0, st>>>(g, Z, ell, var, Lc);
  GPODE_CK(cudaGetLastError());
  GPODE_CK(cholesky(g, Lc, nullptr, info, st));
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
  if (g.df) return df_nu_backward(g, Z, ell, var, u, save, dnu, d_uprior, d_u, d_Z, d_ell, d_var, reinterpret_cast<real*>(ws_f), st);
  const real* Lc = save;
  const real* a = save + static_cast<size_t>(g.Kc) * g.n * g.n;
  const size_t mm = static_cast<size_t>(g.Kc) * g.n * g.n, rhs = static_cast<size_t>(g.Kc) * g.n * g.NR;
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
  const long m2 = static_cast<long>(g.n) * g.n;
  k_build_S<<<dim3(static_cast<unsigned>((m2 + 255) / 256), g.Kc), 256, 0, st>>>(g, u, bb, S);
  GPODE_CK(cudaGetLastError());
  GPODE_CK(trsm<true>(g, g.n, Lc, S, Y, 0, st));                     // Y = Lc^-T S
  GPODE_CK(trsm<true>(g, g.n, Lc, Y, S, 1, st));                     // X^T = Lc^-T Y^T  (X symmetric) -> S
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
