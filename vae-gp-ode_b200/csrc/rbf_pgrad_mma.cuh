// RBF parameter gradients on the warp-level tensor path (mma.sync), sm_100a.
//
//   dnu'_m = sum_e g_e E_em ,   pg_mc = sum_e g_e E_em x_ec ,   E_em = 2^(A_e + H_m + sum_d x_ed G_md)
// for one (sample l, output k); e runs over every state evaluation of the rollout.  Both contractions are
// GEMM-shaped (theta^T = G X^T with K = D; PG = GE X with K = evaluations), the exponential sits between them:
// the structure of a fused attention backward.  Each warp owns 16*MT inducing points (MMA rows); evaluations
// stream through shared memory in batches and are walked 8 at a time (MMA columns):
//   1. theta^T (16 m x 8 e) = G (16 x D) X^T (D x 8): ONE k-step of m16n8k16 with a two-way fp16 split (head + 2^11-scaled
//      remainder, power-of-two block scales per evaluation and per output dimension, see rbf_kernels.cuh): 3 MMAs
//   2. GE = g_e 2^(theta + H_m + A_e) on the C fragment (4 values per lane), MUFU.EX2
//   3. the C fragment IS the A fragment of the second product (rows m, contraction over the 8 evaluations, with
//      the evaluation order permuted consistently in the B operand -- no shuffles): PG (16 m x 8 c) += GE X as
//      TF32 head x head (m16n8k8) + ONE bf16 m16n8k16 for both cross terms rem(GE) X + GE rem(X): GE keeps the fp32 exponent
// Per (16 m x 8 e) tile at D = 16: 7 HMMA + 4 MUFU per lane + ~40 FMA/ALU-pipe instructions, against 140
// FMA-pipe cycles for the same tile on the FFMA path (k_rbf_pgrad): the tensor pipe takes the two dot products.
#pragma once

#include "common.cuh"
#include "rbf.h"

namespace gpode {

constexpr int kPgmThreads = 256;   // 8 warps
constexpr int kPgmBatch = 512;     // evaluations per shared-memory batch (dynamic shared memory: 2 x DK x (batch + 8) floats + offsets)
inline int rbf_pgrad_mma_smem_bytes(int KS) { return (3 * 8 * KS * (kPgmBatch + 8) + 2 * (kPgmBatch + 8) + kPgmBatch + 8 * KS) * 4; }

// TF32 head of x by truncation (one LOP3; cvt.rna.tf32 is emulated with 4 ALU instructions on sm_100a): the remainder
// x - head is exact in fp32 and < 2^-10 |x|; it enters the cross-term MMA rounded to bf16 (~2^-19 relative overall)
__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xFFFFE000u; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// KS: MMA k-steps over the input dimension (1: DP <= 8, 2: DP <= 16); MT: 16-row inducing tiles per warp
template <int KS, int MT>
__global__ void __launch_bounds__(kPgmThreads, (MT == 1 ? 3 : 2)) k_rbf_pgrad_mma(const RbfPgradArgs a) {
  const RbfGeom& g = a.g;
  constexpr int DK = 8 * KS;                 // padded input dimension
  constexpr int NE = kPgmBatch;
  constexpr int TS = NE + 8;                 // row stride of the transposed copies (== 8 mod 32: conflict-free LDS.64)
  // [d][e] TF32 head / remainder of the staged states: B operand of BOTH products (row stride == 8 mod 32 floats makes the
  // per-lane LDS.32 of product 1 (bank 8 tq + gq) and the LDS.64 of product 2 (bank 8 gq + 2 tq per half warp) conflict free)
  extern __shared__ __align__(16) float pgm_smem[];
  constexpr int TSH = NE / 2 + 4;            // row stride of the evaluation-pair copies (== 4 mod 32: conflict-free LDS.32 at gq * TSH + tq)
  float* s_th = pgm_smem;
  // cross-term operand of product 2 (bf16 m16n8k16): [d][evaluation pair] packed bf16 of the states / of their TF32 remainders
  uint32_t* s_pc0 = reinterpret_cast<uint32_t*>(s_th + DK * TS);
  uint32_t* s_pc1 = s_pc0 + DK * TSH;
  float* s_A = reinterpret_cast<float*>(s_pc1 + DK * TSH);   // [NE + 8]: the software-pipelined theta of the block after the last reads 8 unused offsets
  float* s_g = s_A + NE + 8;         // [NE]
  float* s_c = s_g + NE;             // [DK]
  // product 1 runs as fp16 m16n8k16 (two-way split, see rbf_kernels.cuh): the staged states once more as packed fp16 pairs
  // [dim pair][e] (head / 2^11-scaled remainder) of x scaled by a power of two per evaluation, and the un-scaling factor per evaluation
  uint32_t* s_xh2 = reinterpret_cast<uint32_t*>(s_c + DK);
  uint32_t* s_xl2 = s_xh2 + (DK / 2) * TS;
  float* s_u = reinterpret_cast<float*>(s_xl2 + (DK / 2) * TS);   // [NE + 8]
  const int k = blockIdx.y, l = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const float* hdr = rbf_hdr_ptr(a.packed, g, l) + k * g.hdr_floats;
  if (tid < DK) s_c[tid] = tid < g.DP ? hdr[tid] : 0.f;
  const int per_cta = (kPgmThreads / 32) * 16 * MT;
  const int n_mblk = (2 * g.MP2 + per_cta - 1) / per_cta;
  const int chunk_id = blockIdx.x / n_mblk;
  const int m_base = (blockIdx.x - chunk_id * n_mblk) * per_cta + warp * 16 * MT;

  // loop-invariant A fragments of product 1 (fp16 pairs): a[h + 2 j] = sb G[m = gq + 8 h][dims 2 tq + 8 j, + 1], and H_m
  static_assert(KS == 2, "the fp16 form of product 1 covers the 16 padded input dims in one m16n8k16 k-step");
  uint32_t Gh[MT][4], Gl[MT][4];
  float Hm[MT][2];
  float sb, isb;
  pow2_scales(rbf_maxabs_ptr(a.packed, g, l)[k], sb, isb);
  const float* rows = rbf_rows_ptr(a.packed, g, l) + (static_cast<size_t>(k) * (g.SP2 + g.MP2) + g.SP2) * g.row_floats;
  auto packed_at = [&](int m, int q) -> float {   // q < DP: G_md, q == DP: H_m ; rows hold {even m, odd m} float2 pairs
    if (m >= 2 * g.MP2) return 0.f;
    return rows[static_cast<size_t>(m >> 1) * g.row_floats + 2 * q + (m & 1)];
  };
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = m_base + mt * 16 + gq + 8 * h;
      Hm[mt][h] = packed_at(m, g.DP);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int d = 2 * tq + 8 * j;
        const float v0 = d < g.DP ? packed_at(m, d) : 0.f, v1 = d + 1 < g.DP ? packed_at(m, d + 1) : 0.f;
        split_h2(v0 * sb, v1 * sb, Gh[mt][h + 2 * j], Gl[mt][h + 2 * j]);
      }
    }
  }
  float PG[MT][KS][4], dnu[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    dnu[mt][0] = dnu[mt][1] = 0.f;
#pragma unroll
    for (int cb = 0; cb < KS; ++cb)
#pragma unroll
      for (int i = 0; i < 4; ++i) PG[mt][cb][i] = 0.f;
  }

  const long total = a.n_te * g.N;
  const long per = (total + a.chunks - 1) / a.chunks;
  const long e_lo = static_cast<long>(chunk_id) * per;
  const long e_hi = e_lo + per < total ? e_lo + per : total;
  __syncthreads();
  for (long e0 = e_lo; e0 < e_hi; e0 += NE) {
    // ---- stage a batch: thread <-> evaluations tid, tid + 256 (all global loads first, then the splits) ----
    int plain_here = 1;
    {
      constexpr int PER = NE / kPgmThreads;
      static_assert(NE % kPgmThreads == 0, "full warps in the staging shuffles");
      float xv[PER][DK], gv[PER];
      // (te, n) of the first evaluation of the batch by one division; the threads step from there
      const long te0 = e0 / g.N;
      const long n0 = e0 - te0 * g.N;
#pragma unroll
      for (int p = 0; p < PER; ++p) {
        const int idx = tid + p * kPgmThreads;
        const bool ok = e0 + idx < e_hi;
        long te = te0, n = n0 + idx;
        while (n >= g.N) {
          n -= g.N;
          ++te;
        }
        const float* xp = a.xsave + (te * g.D_in) * g.NL + static_cast<long>(l) * g.N + n;
#pragma unroll
        for (int d = 0; d < DK; ++d) xv[p][d] = (ok && d < g.D_in) ? xp[static_cast<long>(d) * g.NL] : 0.f;
        gv[p] = ok ? a.gsave[(te * g.D_out + k) * g.NL + static_cast<long>(l) * g.N + n] : 0.f;     // g = 0 switches padded evaluations off
      }
#pragma unroll
      for (int p = 0; p < PER; ++p) {
        const int idx = tid + p * kPgmThreads;
        float part = 0.f, mx = 0.f;
#pragma unroll
        for (int d = 0; d < DK; ++d) {
          const float v = xv[p][d];
          mx = fmaxf(mx, fabsf(v));
          part = fmaf(s_c[d] * v, v, part);
          const float vh = __uint_as_float(tf32_hi(v)), vl = v - vh;
          s_th[d * TS + idx] = vh;
          const float vn = __shfl_xor_sync(0xffffffffu, v, 1), vln = __shfl_xor_sync(0xffffffffu, vl, 1);   // evaluation idx + 1
          if (!(idx & 1)) {
            s_pc0[d * TSH + (idx >> 1)] = pack_bf16(v, vn);
            s_pc1[d * TSH + (idx >> 1)] = pack_bf16(vl, vln);
          }
        }
        float sa, isa;
        pow2_scales(mx, sa, isa);
#pragma unroll
        for (int dp = 0; dp < DK / 2; ++dp) split_h2(xv[p][2 * dp] * sa, xv[p][2 * dp + 1] * sa, s_xh2[dp * TS + idx], s_xl2[dp * TS + idx]);
        s_u[idx] = isa * isb;
        plain_here = plain_here && (isa * isb == 1.f);
        s_A[idx] = part;
        s_g[idx] = gv[p];
      }
    }
    const int plain = __syncthreads_and(plain_here);   // every scale of this batch is 1 (CTA uniform): offsets as accumulator init
    const int nblk = static_cast<int>(((e_hi - e0 < NE ? e_hi - e0 : NE) + 7) / 8);
    // theta^T of block `eb` for every m tile (3 fp16 MMAs each): th = H_m + A_e + u_e (G' . x'), column n = gq <-> evaluation eb + gq
    auto theta = [&](int eb, float (&th)[MT][4]) {
      const float2 Ae = *reinterpret_cast<const float2*>(s_A + eb + 2 * tq);      // evaluations eb + 2 tq, + 1 (C columns)
      uint32_t bh[2], bl[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        bh[j] = s_xh2[(tq + 4 * j) * TS + eb + gq];     // dims 2 tq + 8 j, + 1 of evaluation eb + gq
        bl[j] = s_xl2[(tq + 4 * j) * TS + eb + gq];
      }
      if (plain) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          float x2[4] = {0.f, 0.f, 0.f, 0.f};
          mma_f16_sweep(x2, Gl[mt], bh[0], bh[1]);
          mma_f16_sweep(x2, Gh[mt], bl[0], bl[1]);
          th[mt][0] = Hm[mt][0] + Ae.x;
          th[mt][1] = Hm[mt][0] + Ae.y;
          th[mt][2] = Hm[mt][1] + Ae.x;
          th[mt][3] = Hm[mt][1] + Ae.y;
          mma_f16_sweep(th[mt], Gh[mt], bh[0], bh[1]);
#pragma unroll
          for (int i = 0; i < 4; ++i) th[mt][i] = fmaf(x2[i], kLoScale, th[mt][i]);
        }
      } else {
        const float2 ue = *reinterpret_cast<const float2*>(s_u + eb + 2 * tq);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          float x2[4] = {0.f, 0.f, 0.f, 0.f}, c[4] = {0.f, 0.f, 0.f, 0.f};
          mma_f16_sweep(x2, Gl[mt], bh[0], bh[1]);
          mma_f16_sweep(x2, Gh[mt], bl[0], bl[1]);
          mma_f16_sweep(c, Gh[mt], bh[0], bh[1]);
          th[mt][0] = fmaf(x2[0], ue.x * kLoScale, fmaf(c[0], ue.x, Hm[mt][0] + Ae.x));
          th[mt][1] = fmaf(x2[1], ue.y * kLoScale, fmaf(c[1], ue.y, Hm[mt][0] + Ae.y));
          th[mt][2] = fmaf(x2[2], ue.x * kLoScale, fmaf(c[2], ue.x, Hm[mt][1] + Ae.x));
          th[mt][3] = fmaf(x2[3], ue.y * kLoScale, fmaf(c[3], ue.y, Hm[mt][1] + Ae.y));
        }
      }
    };
    float thA[MT][4];
    theta(0, thA);
#pragma unroll 1
    for (int cb8 = 0; cb8 < nblk; ++cb8) {
      const int eb = cb8 * 8;
      // software pipeline: the tensor work of the NEXT block is independent of the elementwise work of this one
      // (block nblk reads the 8 pad columns / stale data: its result is never used)
      float thB[MT][4];
      theta(eb + 8, thB);
      const float2 ge = *reinterpret_cast<const float2*>(s_g + eb + 2 * tq);      // evaluations eb + 2 tq, + 1 (C columns)
      // B fragments of product 2: rows = evaluations (MMA k = tq <-> e = 2 tq, k = tq + 4 <-> e = 2 tq + 1), cols c = gq + 8 cb
      // cross terms (bf16 k16): k = 2 tq, 2 tq + 1 <-> x of those evaluations (times the remainders of GE), k + 8 <-> the TF32 remainders of x
      uint32_t xh2[KS][2], xc[KS][2];
#pragma unroll
      for (int cb = 0; cb < KS; ++cb) {
        const float2 vh = *reinterpret_cast<const float2*>(s_th + (gq + 8 * cb) * TS + eb + 2 * tq);
        xh2[cb][0] = __float_as_uint(vh.x); xh2[cb][1] = __float_as_uint(vh.y);
        xc[cb][0] = s_pc0[(gq + 8 * cb) * TSH + (eb >> 1) + tq];
        xc[cb][1] = s_pc1[(gq + 8 * cb) * TSH + (eb >> 1) + tq];
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const float (&th)[4] = thA[mt];
        // C fragment: th[0] (m = gq, e = 2 tq), th[1] (gq, 2 tq + 1), th[2] (gq + 8, 2 tq), th[3] (gq + 8, 2 tq + 1)
        const float ge00 = ge.x * ex2_approx(th[0]), ge01 = ge.y * ex2_approx(th[1]);
        const float ge10 = ge.x * ex2_approx(th[2]), ge11 = ge.y * ex2_approx(th[3]);
        dnu[mt][0] += ge00 + ge01;
        dnu[mt][1] += ge10 + ge11;
        // A fragment of product 2: a0 (row gq, k tq) = ge00, a1 (row gq + 8, k tq) = ge10, a2 (gq, tq + 4) = ge01, a3 = ge11
        uint32_t ah[4], ac[4];
        float al[4];
        const float v4[4] = {ge00, ge10, ge01, ge11};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ah[i] = tf32_hi(v4[i]);
          al[i] = v4[i] - __uint_as_float(ah[i]);
        }
        // cross-term fragment: (row gq | gq + 8, k 2 tq, 2 tq + 1) = remainders of GE, (k + 8) = GE
        ac[0] = pack_bf16(al[0], al[2]);
        ac[1] = pack_bf16(al[1], al[3]);
        ac[2] = pack_bf16(ge00, ge01);
        ac[3] = pack_bf16(ge10, ge11);
#pragma unroll
        for (int cb = 0; cb < KS; ++cb) {
          mma_bf16_sweep(PG[mt][cb], ac, xc[cb][0], xc[cb][1]);
          mma_tf32(PG[mt][cb], ah, xh2[cb][0], xh2[cb][1]);
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) thA[mt][i] = thB[mt][i];
    }
    __syncthreads();
  }
  // ---- flush: PG C fragments (m = gq (+8), c = 2 tq (+1) + 8 cb) and the row sums dnu ----
  const size_t base = (static_cast<size_t>(l) * g.D_out + k) * (2 * g.MP2);
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = m_base + mt * 16 + gq + 8 * h;
      float v = dnu[mt][h];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (m < 2 * g.MP2) {
        if (tq == 0) atomicAdd(&a.acc.dnu[base + m], v);
#pragma unroll
        for (int cb = 0; cb < KS; ++cb)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int c = 8 * cb + 2 * tq + j;
            if (c < g.DP) atomicAdd(&a.acc.pg[(base + m) * g.DP + c], PG[mt][cb][2 * h + j]);
          }
      }
    }
  }
}

}  // namespace gpode
