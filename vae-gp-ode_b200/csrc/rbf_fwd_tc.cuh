// RBF forward sweep on the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory), sm_100a, D > 8.
//
// A CTA owns 256 states = two 128-row MMA tiles (row = TMEM lane = one thread's state).  Per evaluation and output k the
// parameter rows arrive as pre-laid operand tiles of 256 units (k_rbf_pack_tc: no-swizzle K-major core matrices, TF32 head /
// remainder / offset columns, then the 256 weights), one 42 KB bulk copy each, 3-stage ring:
//   theta (128 states x 128 units per MMA) = A B^T over K = 56 in seven kind::tf32 k-steps (3xTF32 along K, fp32 accumulate in TMEM):
//       X_h G_h (2) + X_l G_h (2) + X_h G_l (2) + [1 1 0 0 ..][off_h off_l 0 0 ..] (1)      ~2^-21 relative, fp32 exponent range
//   the epilogue threads then read THEIR state's thetas from tensor memory (tcgen05.ld, 16 columns at a time, the next load in
//   flight) and accumulate
//       f_prior += w cos(theta)      |      f_update += nu' 2^(theta + A_k(x))
// in registers: no operand splits, no fragment shuffles, no cross-lane reduction -- the instruction stream is MUFU + FFMA
// (+ one FADD for A_k), which makes the sweep MUFU-bound.  Tensor memory holds four accumulators of 128 columns (2 state tiles x
// 2 unit halves of a block): the MMAs of one half run under the epilogue of the other (the cost of a tcgen05.mma scales with N
// down to 128 -- 139 -> 69 cycles -- but not to 64).  16 epilogue warps: warp w and warp w + 8 share 32 states, each takes 64 of
// the 128 columns of every accumulator (4 warps per scheduler; with 2 the MUFU pipe stalls at 65 %); the pair's partial sums meet
// through shared memory and an mbarrier once per output dimension.  A 17th warp is producer (bulk copies) and MMA issuer: a
// tcgen05.mma blocks its issuing thread for its duration, so it must not sit inside an epilogue warp.  Completion is tracked with
// tcgen05.commit -> mbarrier, every wait is bounded (a lost arrival traps instead of hanging the GPU).
#pragma once

#include "rbf_kernels.cuh"
#include "tc_common.cuh"   // descriptor / tcgen05 helpers

namespace gpode {

constexpr int kFtStates = 256;                 // states per CTA: threads 0..255 own one state each (solver glue, stores)
constexpr int kFtEpi = 2 * kFtStates;          // epilogue threads: warp w and warp w + 8 share the states of warp w -- of every 128-unit accumulator
                                               // w takes the first 64 columns, w + 8 the second 64 (4 warps per scheduler keep the MUFU pipe fed; 2 reach 65 %)
constexpr int kFtThreads = kFtEpi + 32;        // + one producer / MMA-issuer warp (a tcgen05.mma blocks its issuing thread ~140 cycles)
constexpr int kFtStages = 3;
constexpr int kFtAFloats = kTcfChunks * 128 * 4;     // operand tile of 128 states
constexpr int kFtHalf = 128;                         // units per MMA (N): every state tile has two accumulators of 128 columns -- the MMAs of one
                                                     // run under the epilogue of the other

// tensor-memory load without the wait (16 consecutive columns of this thread's lane) / the wait, which hands the registers over
__device__ __forceinline__ void tc_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
                 "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]),
                 "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}
template <class B>
__device__ __forceinline__ bool tc_try(B bar, uint32_t parity) { return mbar_try(bar, parity); }
__device__ __forceinline__ void lds128(uint32_t saddr, float& a, float& b, float& c, float& d) {
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(saddr));
}

struct FwdTcSmem {
  float* xs;          // [DP][256] staged states (generic solver glue)
  float* hdr;         // [D_out][hdr_floats]
  float* part;        // [2 (k parity)][2 (prior, update)][256]: partial sums of the second-half warps
  float* A;           // 2 x kFtAFloats
  float* B;           // kFtStages x kTcfTileFloats
  uint64_t* bars;     // full[3], empty[3], acc_full[tile], acc_empty[tile]
  uint32_t* tmem_slot;
  const float* tiles; // operand tiles of this sample (global)
  uint32_t tmem;
  uint32_t pair_par;  // phase bits of this warp pair's two hand-over barriers
  long blk;           // running tile counter of this CTA
  long total;
};

inline int rbf_fwd_tc_smem_bytes(const RbfGeom& g) {
  return (16 * kFtThreads + 4 * kFtStates + g.D_out * g.hdr_floats + 2 * kFtAFloats + kFtStages * kTcfTileFloats) * 4 + 384 + 1024;
}

template <int DP_>
struct RbfTcFwdPolicy {
  static constexpr int DP = DP_;
  static constexpr int R = 1;
  static constexpr int kThreads = kFtThreads;
  static constexpr int kMinBlocks = 1;
  static constexpr int kStateThreads = kFtStates;
  static constexpr int kXsStride = 0;       // staging buffers xs / dx are strided by the block size
  static constexpr bool kCoopGlue = true;   // sweep.cuh: the solver glue between evaluations is spread over all 544 threads
  static constexpr int kThreadsBwd = kFtThreads;
  static constexpr int kMinBlocksBwd = 1;
  using Geom = RbfGeom;
  using Accum = RbfAccum;
  using Smem = FwdTcSmem;
  static_assert(DP_ <= 16, "one 16-wide K block per operand part");

  __device__ static __forceinline__ uint64_t* full(const Smem& sm, int s) { return sm.bars + s; }
  __device__ static __forceinline__ uint64_t* empty(const Smem& sm, int s) { return sm.bars + kFtStages + s; }
  __device__ static __forceinline__ uint64_t* acc_full(const Smem& sm, int t, int h) { return sm.bars + 2 * kFtStages + 2 * t + h; }
  __device__ static __forceinline__ uint64_t* acc_empty(const Smem& sm, int t, int h) { return sm.bars + 2 * kFtStages + 4 + 2 * t + h; }
  __device__ static __forceinline__ uint64_t* pair(const Smem& sm, int w, int par) { return sm.bars + 2 * kFtStages + 8 + 2 * w + par; }

  __device__ static __forceinline__ Smem carve(float* smem, const Geom& g) {
    Smem s;
    // (1 KB alignment as an OFFSET to the shared-memory array: a pointer rebuilt from an integer loses its address space and every access
    //  through it becomes a generic LD / ST -- see rbf_bwd_tc.cuh)
    uint32_t pad = (1024u - (smem_u32(smem) & 1023u)) & 1023u;
    asm volatile("" : "+r"(pad));
    float* base = smem + pad / 4;
    s.A = base;
    s.B = s.A + 2 * kFtAFloats;
    s.xs = s.B + kFtStages * kTcfTileFloats;
    s.part = s.xs + 16 * kFtThreads;
    s.hdr = s.part + 4 * kFtStates;
    s.bars = reinterpret_cast<uint64_t*>(s.hdr + g.D_out * g.hdr_floats + ((g.D_out * g.hdr_floats) & 1));
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.bars + 2 * kFtStages + 8 + 16);
    s.blk = 0;
    s.pair_par = 0;
    return s;
  }

  // bulk copy of operand tile `b` of the CTA's sequence (tiles repeat every evaluation) into ring slot b % stages
  __device__ static __forceinline__ void fetch(const Smem& sm, const Geom& g, long b) {
    const int per_eval = g.D_out * rbf_tc_blocks(g);
    const int slot = static_cast<int>(b % kFtStages);
    const float* src = sm.tiles + static_cast<size_t>(b % per_eval) * kTcfTileFloats;
    mbar_expect_tx(full(sm, slot), kTcfTileFloats * 4u);
    bulk_g2s(sm.B + slot * kTcfTileFloats, src, kTcfTileFloats * 4u, full(sm, slot));
  }

  __device__ static __forceinline__ long setup(Smem& sm, ChunkPipe&, const Geom& g, const float* packed, long n_evals, bool) {
    const int l = blockIdx.y, tid = threadIdx.x;
    const float* hdr = rbf_hdr_ptr(packed, g, l);
    for (int i = tid; i < g.D_out * g.hdr_floats; i += blockDim.x) sm.hdr[i] = hdr[i];
    for (int i = tid; i < 16 * kFtThreads; i += blockDim.x) sm.xs[i] = 0.f;
    sm.tiles = rbf_tc_tiles_ptr(packed, g, l);
    sm.total = n_evals * g.D_out * rbf_tc_blocks(g);
    // constant chunks of the state operand: chunk 8 = (1, 1, 0, 0) meets (off_h, off_l, 0, 0); chunk 9 = 0
    if (tid < kFtStates) {
      float* At = sm.A + (tid >> 7) * kFtAFloats;
      const int row = tid & 127;
      *reinterpret_cast<float4*>(At + tc_chunk_off(128, row, 8)) = make_float4(1.f, 1.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(At + tc_chunk_off(128, row, 9)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid == 0) {
      for (int i = 0; i < kFtStages; ++i) {
        mbar_init(full(sm, i), 1);
        mbar_init(empty(sm, i), 1 + kFtEpi / 32);
      }
      for (int t = 0; t < 2; ++t)
        for (int h = 0; h < 2; ++h) {
          mbar_init(acc_full(sm, t, h), 1);
          mbar_init(acc_empty(sm, t, h), 8);
        }
      for (int i = 0; i < 16; ++i) mbar_init(sm.bars + 2 * kFtStages + 8 + i, 1);
      mbar_fence_init();
    }
    if (tid < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    sm.tmem = *sm.tmem_slot;
    if (tid == 0)
      for (long b = 0; b < kFtStages && b < sm.total; ++b) fetch(sm, g, b);
    return sm.total;
  }

  __device__ static __forceinline__ void finish(Smem& sm) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sm.tmem), "r"(512) : "memory");
  }

  // the seven k-steps of state tile `t` against units [128 h, 128 h + 128) of ring slot `slot`, into accumulator half h (one thread)
  __device__ static __forceinline__ void issue_half(const Smem& sm, int t, int slot, int h) {
    const uint32_t a0 = smem_u32(sm.A + t * kFtAFloats), b0 = smem_u32(sm.B + slot * kTcfTileFloats) + h * (kFtHalf / 8) * 128;
    constexpr uint32_t idesc = tc_idesc(128, kFtHalf);
    constexpr int ach[7] = {0, 2, 4, 6, 0, 2, 8}, bch[7] = {0, 2, 0, 2, 4, 6, 8};
#pragma unroll
    for (int s = 0; s < 7; ++s)
      tc_mma_ss(sm.tmem + t * kTcfRows + h * kFtHalf, tc_desc(a0 + ach[s] * 128 * 16, 128 * 16), tc_desc(b0 + bch[s] * kTcfRows * 16, kTcfRows * 16), idesc, s > 0);
  }

  template <class Store>
  __device__ static __forceinline__ void eval_fwd(ChunkPipe&, const Geom& g, long, Smem& sm, Store&& store) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sidx = tid & (kFtStates - 1);          // state slot of this epilogue thread
    const int tile = sidx >> 7, row = sidx & 127, h = (tid >> 8) & 1;   // h: which 64 columns of every accumulator this warp takes
    const int nbs = rbf_tc_blocks_s(g), nb = rbf_tc_blocks(g);
    const int n = g.D_out * nb;   // blocks of one evaluation
    const long b0 = sm.blk;
    if (tid < kFtStates) {   // ---- this state's operand row: TF32 heads and remainders ----
      float* At = sm.A + tile * kFtAFloats;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float xv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = 4 * c + i < DP ? sm.xs[(4 * c + i) * blockDim.x + tid] : 0.f;
        float4 hd, lo_;
        hd.x = __uint_as_float(__float_as_uint(xv[0]) & 0xFFFFE000u);
        hd.y = __uint_as_float(__float_as_uint(xv[1]) & 0xFFFFE000u);
        hd.z = __uint_as_float(__float_as_uint(xv[2]) & 0xFFFFE000u);
        hd.w = __uint_as_float(__float_as_uint(xv[3]) & 0xFFFFE000u);
        lo_ = make_float4(xv[0] - hd.x, xv[1] - hd.y, xv[2] - hd.z, xv[3] - hd.w);
        *reinterpret_cast<float4*>(At + tc_chunk_off(128, row, c)) = hd;
        *reinterpret_cast<float4*>(At + tc_chunk_off(128, row, 4 + c)) = lo_;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid >= kFtEpi) {
      // =============== producer / MMA issuer (one thread): runs ahead of the epilogues as far as the accumulators allow ===============
      // The four accumulators (tile, half) are served in the order they drain (non-blocking polls): no head-of-line blocking
      if (lane == 0) {
        int nxt[4], slot_g[4];
        uint32_t rpar[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          nxt[q] = 0;
          slot_g[q] = static_cast<int>(b0 % kFtStages);
          rpar[q] = static_cast<uint32_t>((b0 / kFtStages) & 1);
        }
        int done = 0, cnt = 0, spins = 0;   // cnt: 4-bit counters of the accumulators issued per ring slot
        unsigned long long idle_t0 = 0;
        while (done < 4) {
          bool any = false;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int tq = q & 1, hq = q >> 1, i = nxt[q];
            if (i >= n) continue;
            const long b = b0 + i;
            const int slot = slot_g[q];
            if (!tc_try(full(sm, slot), rpar[q])) continue;
            if (b >= 1 && !tc_try(acc_empty(sm, tq, hq), static_cast<uint32_t>((b - 1) & 1))) continue;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_half(sm, tq, slot, hq);
            tc_commit(acc_full(sm, tq, hq));
            any = true;
            nxt[q] = i + 1;
            if (i + 1 == n) ++done;
            if (++slot_g[q] == kFtStages) {
              slot_g[q] = 0;
              rpar[q] ^= 1u;
            }
            cnt += 1 << (4 * slot);
            if (((cnt >> (4 * slot)) & 15) == 4) {   // block b is issued on all four accumulators
              cnt &= ~(15 << (4 * slot));
              tc_commit(empty(sm, slot));
              // block b - 1 is drained by every warp (its accumulators were all seen empty): its ring slot takes block b + 2
              if (b >= 1 && b + 2 < sm.total) {
                const int ps = slot == 0 ? kFtStages - 1 : slot - 1;
                tc_wait(empty(sm, ps), static_cast<uint32_t>(((b - 1) / kFtStages) & 1));
                fetch(sm, g, b + 2);
              }
            }
          }
          if (any) spins = 0;
          else if (++spins > 4096) {   // time-bounded like every other wait (common.cuh)
            if (spins == 4097) idle_t0 = global_ns();
            __nanosleep(128);
            if (GPODE_WAIT_TIMEOUT_NS != 0ull && global_ns() - idle_t0 > GPODE_WAIT_TIMEOUT_NS) __trap();
          }
        }
      }
    } else {
      // =============== epilogue warps: thread <-> (state, 64-column share of both accumulators) ===============
      const uint32_t ta0 = sm.tmem + tile * kTcfRows + h * (kFtHalf / 2) + (static_cast<uint32_t>((warp & 3) * 32) << 16);
      int slot = static_cast<int>(b0 % kFtStages);
      uint32_t par = static_cast<uint32_t>(b0 & 1);
      for (int k = 0; k < g.D_out; ++k) {
        const float* hdr_k = sm.hdr + k * g.hdr_floats;
        float Ak = 0.f;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
          const float xv = sm.xs[d * blockDim.x + sidx];
          Ak = fmaf(hdr_k[d] * xv, xv, Ak);
        }
        float fp[1], fu[1];
        float acc0 = 0.f, acc1 = 0.f;
        for (int j = 0; j < nb; ++j) {
          const bool is_k = j >= nbs;
          if (j == nbs) {
            fp[0] = acc0 + acc1;
            acc0 = acc1 = 0.f;
          }
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
          tc_wait(acc_full(sm, tile, hh), par);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t wa = smem_u32(sm.B + slot * kTcfTileFloats + kTcfBFloats + hh * kFtHalf + h * (kFtHalf / 2));
          const uint32_t ta = ta0 + hh * kFtHalf;
          // 16 columns at a time; the tensor-memory load of the next 16 is in flight while these are processed
          uint32_t r[2][16];
          tc_ld16_async(ta, r[0]);
#pragma unroll
          for (int s = 0; s < kFtHalf / 32; ++s) {
            float w[16];
#pragma unroll
            for (int v = 0; v < 4; ++v) lds128(wa + (16 * s + 4 * v) * 4, w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
            tc_ld_wait(r[s & 1]);
            if (s + 1 < kFtHalf / 32) tc_ld16_async(ta + 16 * (s + 1), r[(s + 1) & 1]);
            if (is_k) {
#pragma unroll
              for (int v = 0; v < 16; v += 2) {
                acc0 = fmaf(ex2_approx(__uint_as_float(r[s & 1][v]) + Ak), w[v], acc0);
                acc1 = fmaf(ex2_approx(__uint_as_float(r[s & 1][v + 1]) + Ak), w[v + 1], acc1);
              }
            } else {
#pragma unroll
              for (int v = 0; v < 16; v += 2) {
                acc0 = fmaf(__cosf(__uint_as_float(r[s & 1][v])), w[v], acc0);
                acc1 = fmaf(__cosf(__uint_as_float(r[s & 1][v + 1])), w[v + 1], acc1);
              }
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tc_arrive(acc_empty(sm, tile, hh));
            if (hh == 1) tc_arrive(empty(sm, slot));
          }
          }
          par ^= 1u;
          if (++slot == kFtStages) slot = 0;
        }
        if (nb == nbs) {   // (no inducing section: cannot happen for M >= 1, kept for completeness)
          fp[0] = acc0 + acc1;
          acc0 = acc1 = 0.f;
        }
        fu[0] = acc0 + acc1;
        // the second-half warps hand their partial sums to the state threads: warp w + 8 -> warp w through an mbarrier per
        // (warp pair, k parity); the buffers are double-buffered over k and the accumulator flow control keeps the pair within
        // one output dimension of each other
        float* part = sm.part + (k & 1) * 2 * kFtStates;
        if (h == 1) {
          part[sidx] = fp[0];
          part[kFtStates + sidx] = fu[0];
          __syncwarp();
          if (lane == 0) tc_arrive(pair(sm, warp & 7, k & 1));
        } else {
          tc_wait(pair(sm, warp & 7, k & 1), (sm.pair_par >> (k & 1)) & 1u);
          sm.pair_par ^= 1u << (k & 1);
          fp[0] += part[sidx];
          fu[0] = (fu[0] + part[kFtStates + sidx]) * kInvLn2;   // inducing-row weights carry ln2
          store(k, fp, fu);
        }
      }
    }
    sm.blk = b0 + n;
    __syncthreads();   // every epilogue of this evaluation is done before the next stage overwrites the state operand
  }
};

}  // namespace gpode
