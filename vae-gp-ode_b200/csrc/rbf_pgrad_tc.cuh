// RBF parameter gradients on the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory), sm_100a.
//
//   dnu'_m = sum_e GE_em ,   pg_mc = sum_e GE_em x_ec ,   GE_em = g_e 2^(A_e + H_m + sum_d x_ed G_md)
// for one (sample l, output k) and a block of 128 inducing points (MMA rows = TMEM lanes); e runs over the state
// evaluations of the rollout in tiles of 64 (MMA columns).  Structure of a fused attention backward:
//   1. theta (128 m x 64 e) = A1 B1^T with K = 56: kind::tf32, A1 / B1 in shared memory as no-swizzle K-major core
//      matrices.  3xTF32: operands are split into a TF32 head and an fp32 remainder (the tensor core truncates the
//      remainder to TF32) and laid along K as A1 = [Gh | Gh | Gl | 1 1 Hh Hl], B1 = [Xh | Xl | Xh | Ah Al 1 1], so ONE
//      accumulation chain yields G.x (hi*hi + hi*lo + lo*hi) + H_m + A_e at ~2^-21 relative.
//   2. each thread owns one row m (TMEM lane): tcgen05.ld, GE = g_e 2^theta (MUFU.EX2), row sum -> dnu, split into
//      head / remainder and tcgen05.st back into tensor memory (the head over theta in place).
//   3. PG (128 m x 16 c) += GE X with K = 64 evaluations: A from TENSOR MEMORY (GEh, GEl), B2 = X^T heads / remainders
//      in shared memory; accumulates in tensor memory over every tile of the CTA, read once at the end.
// Warp-specialised, one CTA per SM: 4 producer warps stage tiles into a 3-deep shared-memory ring (register prefetch of
// the next tile's global loads), ONE thread issues every MMA (theta of tile t+2 is issued right after PG of tile t: the
// tensor pipe executes in issue order, so the double-buffered theta/GE columns need no extra barrier), 8 epilogue warps do
// the exponentials.  Completion is tracked with tcgen05.commit -> mbarrier; every wait is bounded (a lost arrival traps).
#pragma once

#include <cstdio>

#include "common.cuh"
#include "rbf.h"

namespace gpode {

constexpr int kTcEpiWarps = 8;     // epilogue warps: warp w works on TMEM lanes 32 (w & 3) .. + 31 (rows m) and the column half w >> 2
constexpr int kTcProdWarps = 4;    // producer warps per group; two groups stage alternate tiles into the shared-memory ring
constexpr int kTcProdGroups = 2;
constexpr int kTcThreads = (kTcEpiWarps + 1 + kTcProdGroups * kTcProdWarps) * 32;   // + 1 MMA-issuer warp
constexpr int kTcTile = 64;        // evaluations per tile (MMA N of product 1, K of product 2)
constexpr int kTcK1 = 56;          // 16 (hi*hi) + 16 (hi*lo) + 16 (lo*hi) + 4 (offsets) + 4 (pad)
constexpr int kTcCols = 512;       // TMEM columns: 2 x (theta/GEh 64 | GEl 64), PG 32 (one CTA per SM)
constexpr int kTcStages = 4;       // shared-memory ring
constexpr int kTcDim = 16;         // padded input dimension / PG columns

// no-swizzle K-major operand tile [ROWS x K] (fp32 words): 16-byte chunks of 4 consecutive k; chunk c of row r at
// c * ROWS * 4 + (r / 8) * 32 + (r % 8) * 4  ->  LBO = ROWS * 16 bytes (next chunk), SBO = 128 bytes (next 8 rows)
__device__ __forceinline__ int tc_chunk_off(int rows, int row, int chunk) { return chunk * rows * 4 + (row >> 3) * 32 + (row & 7) * 4; }
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return static_cast<uint64_t>((saddr & 0x3FFFF) >> 4) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
         (static_cast<uint64_t>(128 >> 4) << 32) | (static_cast<uint64_t>(1) << 46);
}
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {   // f32 accumulate, tf32 x tf32, both K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity) {   // bounded: a lost completion traps instead of hanging the GPU
  for (int spin = 0; spin < (1 << 26); ++spin) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
                 "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ float tc_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// shared-memory carve-up (floats)
constexpr int kTcA1 = 128 * kTcK1;               // A1: 128 rows x 56
constexpr int kTcB1 = kTcTile * kTcK1;           // B1: 64 rows x 56
constexpr int kTcB2 = 2 * kTcDim * kTcTile;      // B2: 32 rows ([X head ; X remainder]) x 64 evaluations
constexpr int kTcScr = kTcTile * (kTcDim + 1);   // row-major scratch of the staged states
constexpr int kTcStageFloats = kTcB1 + kTcB2 + kTcScr + 2 * kTcTile + kTcTile;   // + partial A_e [2][64] + g_e [64]
constexpr int kTcSmemFloats = kTcA1 + kTcStages * kTcStageFloats + kTcDim;
inline int rbf_pgrad_tc_smem_bytes() { return kTcSmemFloats * 4 + 1024; }

__device__ __forceinline__ void tc_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

template <int kUnused>   // (template only for vague linkage: the header is included by one translation unit per DP)
__global__ void __launch_bounds__(kTcThreads, 1) k_rbf_pgrad_tc(const RbfPgradArgs a) {
  const RbfGeom& g = a.g;
  extern __shared__ __align__(1024) float tc_smem_raw[];
  float* sm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* sA1 = sm;
  float* sStage = sA1 + kTcA1;
  float* sC = sStage + kTcStages * kTcStageFloats;   // [16] c_kd
  __shared__ __align__(8) uint64_t bar_full[kTcStages], bar_empty[kTcStages], bar_th[2], bar_ge[2], bar_done;
  __shared__ uint32_t tmem_base_s;

  const int k = blockIdx.y, l = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_mblk = (2 * g.MP2 + 127) / 128;
  const int chunk_id = blockIdx.x / n_mblk;
  const int m_blk0 = (blockIdx.x - chunk_id * n_mblk) * 128;
  const float* hdr = rbf_hdr_ptr(a.packed, g, l) + k * g.hdr_floats;
  if (tid < kTcDim) sC[tid] = tid < g.DP ? hdr[tid] : 0.f;

  // ---- A1 (loop invariant): row r = [Gh | Gh | Gl | 1 1 Hh Hl | 0] ----
  if (tid < 128) {
    const int mm = m_blk0 + tid;
    const float* rows = rbf_rows_ptr(a.packed, g, l) + (static_cast<size_t>(k) * (g.SP2 + g.MP2) + g.SP2) * g.row_floats;
    const bool real = mm < 2 * g.MP2;
    const float* prow = rows + static_cast<size_t>(mm >> 1) * g.row_floats + (mm & 1);
    float Gv[kTcDim];
#pragma unroll
    for (int d = 0; d < kTcDim; ++d) Gv[d] = (real && d < g.DP) ? prow[2 * d] : 0.f;
    const float H = real ? prow[2 * g.DP] : 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float4 h, lo_;
      h.x = tc_hi(Gv[4 * c]); h.y = tc_hi(Gv[4 * c + 1]); h.z = tc_hi(Gv[4 * c + 2]); h.w = tc_hi(Gv[4 * c + 3]);
      lo_.x = Gv[4 * c] - h.x; lo_.y = Gv[4 * c + 1] - h.y; lo_.z = Gv[4 * c + 2] - h.z; lo_.w = Gv[4 * c + 3] - h.w;
      *reinterpret_cast<float4*>(sA1 + tc_chunk_off(128, tid, c)) = h;
      *reinterpret_cast<float4*>(sA1 + tc_chunk_off(128, tid, 4 + c)) = h;
      *reinterpret_cast<float4*>(sA1 + tc_chunk_off(128, tid, 8 + c)) = lo_;
    }
    const float Hh = tc_hi(H);
    *reinterpret_cast<float4*>(sA1 + tc_chunk_off(128, tid, 12)) = make_float4(1.f, 1.f, Hh, H - Hh);
    *reinterpret_cast<float4*>(sA1 + tc_chunk_off(128, tid, 13)) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (tid == 0) {
    for (int i = 0; i < kTcStages; ++i) {
      mbar_init(&bar_full[i], kTcProdWarps * 32);
      mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_th[i], 1);
      mbar_init(&bar_ge[i], kTcEpiWarps * 32);
    }
    mbar_init(&bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kTcCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t t_pg = tmem + 256;

  const long total = a.n_te * g.N;
  const long per = (total + a.chunks - 1) / a.chunks;
  const long e_lo = static_cast<long>(chunk_id) * per;
  const long e_hi = e_lo + per < total ? e_lo + per : total;
  const int nt = e_hi > e_lo ? static_cast<int>((e_hi - e_lo + kTcTile - 1) / kTcTile) : 0;

  if (warp < kTcEpiWarps) {
    // =========================== epilogue: GE = g_e 2^theta, row sums, split, back to tensor memory ===========================
    const int row = (warp & 3) * 32 + lane, chalf = warp >> 2;
    const int m = m_blk0 + row;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
#ifdef GPODE_TC_TIMERS
    long long tw = 0, tc = 0, t0_;
#endif
    float dnu = 0.f;
    for (int t = 0; t < nt; ++t) {
#ifdef GPODE_TC_TIMERS
      t0_ = clock64();
#endif
      const int b = t & 1, st = t % kTcStages;
      const float* sG = sStage + st * kTcStageFloats + kTcB1 + kTcB2 + kTcScr + 2 * kTcTile;
      tc_wait(&bar_full[st], (t / kTcStages) & 1);          // g_e of this tile is staged (already true when theta is done)
      tc_wait(&bar_th[b], (t >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef GPODE_TC_TIMERS
      { const long long n_ = clock64(); tw += n_ - t0_; t0_ = n_; }
#endif
      const uint32_t t_th = tmem + b * 128, t_gl = t_th + kTcTile;
#pragma unroll 1
      for (int c0 = 32 * chalf; c0 < 32 * chalf + 32; c0 += 16) {
        uint32_t r[16], rl[16];
        tc_ld16(t_th + lane_base + c0, r);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v = sG[c0 + j] * ex2_approx(__uint_as_float(r[j]));
          dnu += v;
          const float h = tc_hi(v);
          r[j] = __float_as_uint(h);
          rl[j] = __float_as_uint(v - h);
        }
        tc_st16(t_th + lane_base + c0, r);
        tc_st16(t_gl + lane_base + c0, rl);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      tc_arrive(&bar_ge[b]);
#ifdef GPODE_TC_TIMERS
      tc += clock64() - t0_;
#endif
    }
#ifdef GPODE_TC_TIMERS
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && nt > 0) printf("tc epilogue: wait %lld compute %lld cycles/tile (%d tiles)\n", tw / nt, tc / nt, nt);
#endif
    if (nt > 0) {
      tc_wait(&bar_done, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[16];
      tc_ld16(t_pg + lane_base + 16 * chalf, r);     // PG = D[:, 0:16] + D[:, 16:32]; each column half flushes one of the two
      if (m < 2 * g.MP2) {
        const size_t base = (static_cast<size_t>(l) * g.D_out + k) * (2 * g.MP2) + m;
        atomicAdd(&a.acc.dnu[base], dnu);            // both column halves add their part of the row sum
#pragma unroll
        for (int c = 0; c < kTcDim; ++c)
          if (c < g.DP) atomicAdd(&a.acc.pg[base * g.DP + c], __uint_as_float(r[c]));
      }
    }
  } else if (warp == kTcEpiWarps) {
    // =========================== MMA issuer (one thread) ===========================
    if (lane == 0 && nt > 0) {
      constexpr uint32_t idesc1 = tc_idesc(128, kTcTile), idesc2 = tc_idesc(128, 2 * kTcDim);
      auto issue_theta = [&](int t) {
        const int st = t % kTcStages, b = t & 1;
        tc_wait(&bar_full[st], (t / kTcStages) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sB1 = smem_u32(sStage + st * kTcStageFloats);
#pragma unroll
        for (int ks = 0; ks < kTcK1 / 8; ++ks)
          tc_mma_ss(tmem + b * 128, tc_desc(smem_u32(sA1) + ks * 2 * (128 * 16), 128 * 16), tc_desc(sB1 + ks * 2 * (kTcTile * 16), kTcTile * 16), idesc1, ks > 0);
        tc_commit(&bar_th[b]);
      };
      issue_theta(0);
      if (nt > 1) issue_theta(1);
#ifdef GPODE_TC_TIMERS
      long long mw = 0, mp = 0, mt = 0, m0_;
#endif
      for (int t = 0; t < nt; ++t) {
#ifdef GPODE_TC_TIMERS
        m0_ = clock64();
#endif
        const int st = t % kTcStages, b = t & 1;
        tc_wait(&bar_ge[b], (t >> 1) & 1);
#ifdef GPODE_TC_TIMERS
        { const long long n_ = clock64(); mw += n_ - m0_; m0_ = n_; }
#endif
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sB2 = smem_u32(sStage + st * kTcStageFloats + kTcB1);
#pragma unroll
        for (int p = 0; p < 2; ++p)        // A = GEh, then GEl (tensor memory); B2 = [Xh ; Xl] (32 rows)
#pragma unroll
          for (int ks = 0; ks < kTcTile / 8; ++ks)
            tc_mma_ts(t_pg, tmem + b * 128 + p * kTcTile + ks * 8, tc_desc(sB2 + ks * 2 * (2 * kTcDim * 16), 2 * kTcDim * 16), idesc2, (t > 0 || p > 0 || ks > 0) ? 1u : 0u);
        tc_commit(&bar_empty[st]);          // stage st (B1, B2, g) and the GE columns of buffer b are free once these MMAs are done
#ifdef GPODE_TC_TIMERS
        { const long long n_ = clock64(); mp += n_ - m0_; m0_ = n_; }
#endif
        if (t + 2 < nt) issue_theta(t + 2);
#ifdef GPODE_TC_TIMERS
        mt += clock64() - m0_;
#endif
      }
#ifdef GPODE_TC_TIMERS
      if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) printf("tc mma: wait_ge %lld pg_issue %lld theta(wait_full+issue) %lld cycles/tile\n", mw / nt, mp / nt, mt / nt);
#endif
      tc_commit(&bar_done);
    }
  } else {
    // =========================== producers: stage tile t into ring slot t % kTcStages ===========================
    const int pgrp = (warp - (kTcEpiWarps + 1)) / kTcProdWarps;                       // this group stages tiles t == pgrp (mod 2)
    const int ptid = tid - (kTcEpiWarps + 1 + pgrp * kTcProdWarps) * 32;              // 0..127 inside the group
    const int idx = ptid & (kTcTile - 1), half = ptid >> 6;  // evaluation of the tile, input dims 8 half .. 8 half + 7
    float xn[8], gn;
    auto prefetch = [&](long e0) {
      const long e = e0 + idx;
      const bool ok = e < e_hi;
      long te = 0, s = 0;
      if (ok) {
        te = e / g.N;
        s = static_cast<long>(l) * g.N + (e - te * g.N);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int d = 8 * half + j;
        xn[j] = (ok && d < g.D_in) ? a.xsave[(te * g.D_in + d) * g.NL + s] : 0.f;
      }
      gn = (ok && half == 0) ? a.gsave[(te * g.D_out + k) * g.NL + s] : 0.f;   // g = 0 switches padded evaluations off
    };
    if (pgrp < nt) prefetch(e_lo + static_cast<long>(pgrp) * kTcTile);
    for (int t = pgrp; t < nt; t += kTcProdGroups) {
      const int st = t % kTcStages;
      float* sB1 = sStage + st * kTcStageFloats;
      float* sB2 = sB1 + kTcB1;
      float* sScr = sB2 + kTcB2;
      float* sPart = sScr + kTcScr;
      float* sG = sPart + 2 * kTcTile;
      float xv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) xv[j] = xn[j];
      const float gv = gn;
      if (t + kTcProdGroups < nt) prefetch(e_lo + static_cast<long>(t + kTcProdGroups) * kTcTile);   // lands while this tile is staged / the ring is full
      if (t >= kTcStages) tc_wait(&bar_empty[st], ((t / kTcStages) - 1) & 1);
      // pass 1: thread <-> (evaluation, 8 input dims): B1 chunks (head, remainder, head), row-major scratch
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int d = 8 * half + j;
        part = fmaf(sC[d] * xv[j], xv[j], part);
        sScr[idx * (kTcDim + 1) + d] = xv[j];
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float4 h, lo_;
        h.x = tc_hi(xv[4 * c]); h.y = tc_hi(xv[4 * c + 1]); h.z = tc_hi(xv[4 * c + 2]); h.w = tc_hi(xv[4 * c + 3]);
        lo_.x = xv[4 * c] - h.x; lo_.y = xv[4 * c + 1] - h.y; lo_.z = xv[4 * c + 2] - h.z; lo_.w = xv[4 * c + 3] - h.w;
        const int ch = 2 * half + c;
        *reinterpret_cast<float4*>(sB1 + tc_chunk_off(kTcTile, idx, ch)) = h;
        *reinterpret_cast<float4*>(sB1 + tc_chunk_off(kTcTile, idx, 4 + ch)) = lo_;
        *reinterpret_cast<float4*>(sB1 + tc_chunk_off(kTcTile, idx, 8 + ch)) = h;
      }
      sPart[half * kTcTile + idx] = part;
      if (half == 0) sG[idx] = gv;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + pgrp), "n"(kTcProdWarps * 32) : "memory");
      // pass 2: thread <-> (input dim c, evaluation quads): B2 = [X head ; X remainder] chunks; offsets chunk of B1
      {
        const int c = ptid & 15;
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          const int q = (ptid >> 4) + 8 * rep;
          float4 h, lo_;
          const float v0 = sScr[(4 * q) * (kTcDim + 1) + c], v1 = sScr[(4 * q + 1) * (kTcDim + 1) + c];
          const float v2 = sScr[(4 * q + 2) * (kTcDim + 1) + c], v3 = sScr[(4 * q + 3) * (kTcDim + 1) + c];
          h.x = tc_hi(v0); h.y = tc_hi(v1); h.z = tc_hi(v2); h.w = tc_hi(v3);
          lo_.x = v0 - h.x; lo_.y = v1 - h.y; lo_.z = v2 - h.z; lo_.w = v3 - h.w;
          *reinterpret_cast<float4*>(sB2 + tc_chunk_off(2 * kTcDim, c, q)) = h;
          *reinterpret_cast<float4*>(sB2 + tc_chunk_off(2 * kTcDim, kTcDim + c, q)) = lo_;
        }
        if (ptid < kTcTile) {
          const float A = sPart[ptid] + sPart[kTcTile + ptid];
          const float Ah = tc_hi(A);
          *reinterpret_cast<float4*>(sB1 + tc_chunk_off(kTcTile, ptid, 12)) = make_float4(Ah, A - Ah, 1.f, 1.f);
          *reinterpret_cast<float4*>(sB1 + tc_chunk_off(kTcTile, ptid, 13)) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor (async) proxy
      tc_arrive(&bar_full[st]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcCols) : "memory");
}

}  // namespace gpode
