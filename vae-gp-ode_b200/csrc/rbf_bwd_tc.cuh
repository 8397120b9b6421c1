// RBF reverse sweep on the 5th-generation tensor cores (tcgen05.mma, accumulators AND the second product's A operand in tensor memory),
// sm_100a, D > 8.  Structure of a fused attention backward (S = Q K^T -> P in place -> dQ = P V):
//
// A CTA owns 128 states (row = TMEM lane = one state).  Per evaluation and output k the parameter rows arrive as operand tiles of
// 128 units ("items", k_rbf_pack_tcb: layout in rbf.h), the theta operand and the second-product operand through separate bulk-copy
// rings (their lifetimes differ by three items):
//   1. theta (128 states x 128 units) = A B^T over K = 56, seven kind::tf32 k-steps (3xTF32 along K, exactly the forward's), into one of
//      three 128-column accumulators;
//   2. 16 epilogue warps (warp w: lanes 32 (w & 3).., columns 32 (w >> 2)..) read theta (tcgen05.ld), evaluate
//          tau = -sin(theta)  (feature units; the + pi/2 is in the packed offset)      tau = 2^(theta + A_k(x))  (inducing units)
//      split tau into a bf16 head and a bf16 remainder (16 mantissa bits, fp32 exponent range) and store the PAIR back over theta
//      (tcgen05.st): one 32-bit column = two consecutive K elements of a kind::f16 A operand;
//   3. Q_k (128 states x 48) += tau B2 with A read from TENSOR MEMORY (16 k-steps of kind::f16, K = 16 = 8 units x {head, remainder});
//      B2 = [P_h | w_h || P_l | w_l] with P = weight x coefficient, so the head columns carry (tau_h + tau_l) P_h, the remainder columns
//      tau_h P_l, and column 16 / 40 the weighted sum of tau that the A_k(x) term needs;
//   4. once per k the warps read Q_k: dx_k = g_k (Q + 2 c_d x_d Es), dx += dx_k, lengthscale statistic sum_n x_d dx_kd (one warp reduction).
// The legacy mma.sync kernel spends 7 HMMA + the operand splits per (16 x 8) tile in the instruction stream of the warps that also run
// the transcendentals; here the epilogue stream is LDTM + MUFU + 3 ALU + STTM per element and the products run asynchronously:
// per item the tensor pipe needs 452 + 800 cycles (tools/tc_probe2.cu: a TS kind::f16 MMA costs ~50 cycles whatever N -- it is bound by
// reading the 4 KB A operand from tensor memory), the MUFU pipe 1,024.  A 17th warp issues every MMA and bulk copy; flow control is
// mbarrier + tcgen05.commit, all waits time-bounded.  The tensor pipe executes in issue order, so theta(i + 3) overwriting the accumulator
// of item i is safe once Q(i) has been issued.
#pragma once

#include <cuda_bf16.h>

#include "rbf_kernels.cuh"
#include "tc_common.cuh"

namespace gpode {

constexpr int kBtStates = 128;                 // states per CTA
constexpr int kBtEpiWarps = 16;
constexpr int kBtEpi = kBtEpiWarps * 32;
constexpr int kBtThreads = kBtEpi + 64;        // + MMA-issuer warp + bulk-copy producer warp (issuing a 20 KB copy stalls its thread ~800 cycles)
constexpr int kBtThStages = 3;                 // theta-operand ring (freed when theta(i) has executed)
constexpr int kBtPStages = 4;                  // second-operand ring (freed when Q(i) has executed)
constexpr int kBtAFloats = kTcfChunks * kBtStates * 4;
constexpr int kBtQCol = 3 * kTcbUnits;         // tensor-memory columns: 3 theta accumulators, then 2 Q buffers of 64
constexpr int kBtNBars = 2 * kBtThStages + 2 * kBtPStages + 3 + 3 + 2 + 2;

__host__ __device__ constexpr uint32_t tc_idesc_bf16(int M, int N) {   // f32 accumulate, bf16 x bf16, both K-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_ts_bf16(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tc_st16_async(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ void tc_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3)::"memory");
  v[0] = __uint_as_float(r0);
  v[1] = __uint_as_float(r1);
  v[2] = __uint_as_float(r2);
  v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ float tc_ld1(uint32_t taddr) {
  uint32_t r0;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r0)::"memory");
  return __uint_as_float(r0);
}
// tau -> packed {bf16 head (bits 0..15: the even K element), bf16 remainder (bits 16..31)}: head by truncation (exact remainder), the
// remainder rounded -- |tau - (h + l)| <= 2^-17 |tau|
__device__ __forceinline__ uint32_t tc_split_bf16(float tau) {
  const float h = __uint_as_float(__float_as_uint(tau) & 0xFFFF0000u);
  uint32_t out;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(out) : "f"(tau - h), "f"(h));
  return out;
}

struct BwdTcSmem {
  float* xs;          // [DP][128] staged states (generic solver glue), stride kBtStates
  float* dx;          // [DP][128] J^T g of this evaluation
  float* hdr;         // [D_out][hdr_floats]
  float* dell;        // [D_out][DP] lengthscale statistic, then dvar [D_out] (contiguous, like SweepSmem)
  float* dvar;
  float* A;           // state operand of theta (kBtAFloats)
  float* Bth;         // kBtThStages x kTcbThFloats
  float* Bp;          // kBtPStages x kTcbPFloats
  uint64_t* bars;
  uint32_t* tmem_slot;
  long* ring;         // issuer-only state: tiles fetched so far {theta, second operand}
  const float* tiles; // operand tiles of this sample (global)
  uint32_t tmem;
  long blk;           // running item counter of this CTA
  long kk;            // running (evaluation, k) counter: Q buffer kk & 1
  long total;
};

inline int rbf_bwd_tc_smem_bytes(const RbfGeom& g) {
  return (kBtAFloats + kBtThStages * kTcbThFloats + kBtPStages * kTcbPFloats + 2 * 16 * kBtStates + g.D_out * g.hdr_floats + g.D_out * (g.DP + 1) + 8) * 4 +
         kBtNBars * 8 + 64 + 1024;
}

template <int DP_>
struct RbfTcBwdPolicy {
  static constexpr int DP = DP_;
  static constexpr int R = 1;
  static constexpr int kThreads = kBtThreads;
  static constexpr int kMinBlocks = 1;
  static constexpr int kStateThreads = kBtStates;
  static constexpr int kXsStride = kBtStates;
  static constexpr int kThreadsBwd = kBtThreads;
  static constexpr int kMinBlocksBwd = 1;
  using Geom = RbfGeom;
  using Accum = RbfAccum;
  using Smem = BwdTcSmem;
  static_assert(DP_ <= 16, "one 16-wide K block per operand part");

  __device__ static __forceinline__ uint64_t* th_full(const Smem& sm, int s) { return sm.bars + s; }
  __device__ static __forceinline__ uint64_t* th_empty(const Smem& sm, int s) { return sm.bars + kBtThStages + s; }
  __device__ static __forceinline__ uint64_t* p_full(const Smem& sm, int s) { return sm.bars + 2 * kBtThStages + s; }
  __device__ static __forceinline__ uint64_t* p_empty(const Smem& sm, int s) { return sm.bars + 2 * kBtThStages + kBtPStages + s; }
  __device__ static __forceinline__ uint64_t* acc_full(const Smem& sm, int a) { return sm.bars + 2 * kBtThStages + 2 * kBtPStages + a; }
  __device__ static __forceinline__ uint64_t* tau_ready(const Smem& sm, int a) { return sm.bars + 2 * kBtThStages + 2 * kBtPStages + 3 + a; }
  __device__ static __forceinline__ uint64_t* q_full(const Smem& sm, int q) { return sm.bars + 2 * kBtThStages + 2 * kBtPStages + 6 + q; }
  __device__ static __forceinline__ uint64_t* q_empty(const Smem& sm, int q) { return sm.bars + 2 * kBtThStages + 2 * kBtPStages + 8 + q; }

  __device__ static __forceinline__ Smem carve(float* smem, const Geom& g) {
    Smem s;
    float* base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~static_cast<uintptr_t>(1023));
    s.A = base;
    s.Bth = s.A + kBtAFloats;
    s.Bp = s.Bth + kBtThStages * kTcbThFloats;
    s.xs = s.Bp + kBtPStages * kTcbPFloats;
    s.dx = s.xs + 16 * kBtStates;
    s.hdr = s.dx + 16 * kBtStates;
    s.dell = s.hdr + g.D_out * g.hdr_floats;
    s.dvar = s.dell + g.D_out * DP;
    float* end = s.dvar + g.D_out;
    s.bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(end + 1) + 7) & ~static_cast<uintptr_t>(7));
    s.ring = reinterpret_cast<long*>(s.bars + kBtNBars);
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.ring + 2);
    s.blk = 0;
    s.kk = 0;
    return s;
  }

  __device__ static __forceinline__ void fetch_th(const Smem& sm, const Geom& g, long b) {
    const int per_eval = g.D_out * rbf_tcb_items(g);
    const int slot = static_cast<int>(b % kBtThStages);
    const float* src = sm.tiles + static_cast<size_t>(b % per_eval) * kTcbTileFloats;
    mbar_expect_tx(th_full(sm, slot), kTcbThFloats * 4u);
    bulk_g2s(sm.Bth + slot * kTcbThFloats, src, kTcbThFloats * 4u, th_full(sm, slot));
  }
  __device__ static __forceinline__ void fetch_p(const Smem& sm, const Geom& g, long b) {
    const int per_eval = g.D_out * rbf_tcb_items(g);
    const int slot = static_cast<int>(b % kBtPStages);
    const float* src = sm.tiles + static_cast<size_t>(b % per_eval) * kTcbTileFloats + kTcbThFloats;
    mbar_expect_tx(p_full(sm, slot), kTcbPFloats * 4u);
    bulk_g2s(sm.Bp + slot * kTcbPFloats, src, kTcbPFloats * 4u, p_full(sm, slot));
  }

  __device__ static __forceinline__ long setup(Smem& sm, ChunkPipe&, const Geom& g, const float* packed, long n_evals, bool) {
    const int l = blockIdx.y, tid = threadIdx.x;
    const float* hdr = rbf_hdr_ptr(packed, g, l);
    for (int i = tid; i < g.D_out * g.hdr_floats; i += blockDim.x) sm.hdr[i] = hdr[i];
    for (int i = tid; i < 2 * 16 * kBtStates; i += blockDim.x) sm.xs[i] = 0.f;   // xs and dx
    for (int i = tid; i < g.D_out * (DP + 1); i += blockDim.x) sm.dell[i] = 0.f;  // dell and dvar
    sm.tiles = rbf_tcb_tiles_ptr(packed, g, l);
    sm.total = n_evals * g.D_out * rbf_tcb_items(g);
    if (tid < kBtStates) {   // constant chunks of the state operand: chunk 8 = (1, 1, 0, 0) meets (off_h, off_l, 0, 0); chunk 9 = 0
      *reinterpret_cast<float4*>(sm.A + tc_chunk_off(128, tid, 8)) = make_float4(1.f, 1.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(sm.A + tc_chunk_off(128, tid, 9)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid == 0) {
      for (int i = 0; i < kBtThStages; ++i) {
        mbar_init(th_full(sm, i), 1);
        mbar_init(th_empty(sm, i), 1);
      }
      for (int i = 0; i < kBtPStages; ++i) {
        mbar_init(p_full(sm, i), 1);
        mbar_init(p_empty(sm, i), 1);
      }
      for (int i = 0; i < 3; ++i) {
        mbar_init(acc_full(sm, i), 1);
        mbar_init(tau_ready(sm, i), kBtEpiWarps);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(q_full(sm, i), 1);
        mbar_init(q_empty(sm, i), kBtEpiWarps);
      }
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
#ifdef GPODE_DEBUG_WAIT
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) printf("bwd tc: bars at smem+%u (th_full 0.., th_empty %d.., p_full %d.., p_empty %d.., acc_full %d.., tau_ready %d.., q_full %d.., q_empty %d..) total %ld\n", smem_u32(sm.bars), kBtThStages, 2 * kBtThStages, 2 * kBtThStages + kBtPStages, 2 * kBtThStages + 2 * kBtPStages, 2 * kBtThStages + 2 * kBtPStages + 3, 2 * kBtThStages + 2 * kBtPStages + 6, 2 * kBtThStages + 2 * kBtPStages + 8, sm.total);
#endif
    if (tid == kBtEpi + 32) {   // the producer thread owns the rings: first tiles of both
      long f = 0;
      for (; f < kBtThStages && f < sm.total; ++f) fetch_th(sm, g, f);
      sm.ring[0] = f;
      for (f = 0; f < kBtPStages && f < sm.total; ++f) fetch_p(sm, g, f);
      sm.ring[1] = f;
    }
    return sm.total;
  }

  __device__ static __forceinline__ void finish(Smem&) {}

  __device__ static __forceinline__ void flush(const Smem& sm, const Geom& g, const Accum& acc) {
    for (int i = threadIdx.x; i < g.D_out * DP; i += blockDim.x) atomicAdd(&acc.dell_x[i], sm.dell[i]);
    for (int i = threadIdx.x; i < g.D_out; i += blockDim.x) atomicAdd(&acc.dvar[i], sm.dvar[i]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sm.tmem), "r"(512) : "memory");
  }

  // the issuer is also the producer: before it waits for tile b it makes sure the copy of tile b has been issued (blocking on the slot's
  // previous occupant if the non-blocking prefetch of refill() has not got there yet) -- otherwise it would wait for itself
  __device__ static __forceinline__ void ensure_th(const Smem& sm, const Geom& g, long b) {
    long f = sm.ring[0];
    for (; f <= b; ++f) {
      tc_wait(th_empty(sm, static_cast<int>(f % kBtThStages)), static_cast<uint32_t>(((f - kBtThStages) / kBtThStages) & 1));
      fetch_th(sm, g, f);
    }
    sm.ring[0] = f;
  }
  __device__ static __forceinline__ void ensure_p(const Smem& sm, const Geom& g, long b) {
    long f = sm.ring[1];
    for (; f <= b; ++f) {
      tc_wait(p_empty(sm, static_cast<int>(f % kBtPStages)), static_cast<uint32_t>(((f - kBtPStages) / kBtPStages) & 1));
      fetch_p(sm, g, f);
    }
    sm.ring[1] = f;
  }
  // theta of item b (one thread): seven k-steps into accumulator b % 3; frees the ring slot when executed
  __device__ static __forceinline__ void issue_theta(const Smem& sm, const Geom&, long b) {
    const int slot = static_cast<int>(b % kBtThStages), acc = static_cast<int>(b % 3);
    tc_wait(th_full(sm, slot), static_cast<uint32_t>((b / kBtThStages) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t a0 = smem_u32(sm.A), b0 = smem_u32(sm.Bth + slot * kTcbThFloats);
    constexpr uint32_t idesc = tc_idesc(128, kTcbUnits);
    constexpr int ach[7] = {0, 2, 4, 6, 0, 2, 8}, bch[7] = {0, 2, 0, 2, 4, 6, 8};
#pragma unroll
    for (int s = 0; s < 7; ++s)
      tc_mma_ss(sm.tmem + acc * kTcbUnits, tc_desc(a0 + ach[s] * 128 * 16, 128 * 16), tc_desc(b0 + bch[s] * kTcbUnits * 16, kTcbUnits * 16), idesc, s > 0);
    tc_commit(acc_full(sm, acc));
    tc_commit(th_empty(sm, slot));
  }
  // refill whatever ring slots have drained (non-blocking: the issuer never waits for a copy it does not need yet)
  __device__ static __forceinline__ void refill(const Smem& sm, const Geom& g) {
    long f = sm.ring[0];
    while (f < sm.total && mbar_test(th_empty(sm, static_cast<int>(f % kBtThStages)), static_cast<uint32_t>(((f - kBtThStages) / kBtThStages) & 1))) {
      fetch_th(sm, g, f);
      ++f;
    }
    sm.ring[0] = f;
    f = sm.ring[1];
    while (f < sm.total && mbar_test(p_empty(sm, static_cast<int>(f % kBtPStages)), static_cast<uint32_t>(((f - kBtPStages) / kBtPStages) & 1))) {
      fetch_p(sm, g, f);
      ++f;
    }
    sm.ring[1] = f;
  }

  __device__ static __forceinline__ void vjp(ChunkPipe&, const Geom& g, long, Smem& sm, const States<R>& st, const float* gvec, const float* fvec,
                                             const float* fpvec, long kstride, long sstride) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbs = rbf_tcb_items_s(g), nbi = rbf_tcb_items(g);
    const int n = g.D_out * nbi;   // items of one evaluation
    const long b0 = sm.blk, kk0 = sm.kk;
    if (tid < kBtStates) {   // ---- this state's operand row: TF32 heads and remainders ----
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float xv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = 4 * c + i < DP ? sm.xs[(4 * c + i) * kBtStates + tid] : 0.f;
        float4 hd, lo_;
        hd.x = __uint_as_float(__float_as_uint(xv[0]) & 0xFFFFE000u);
        hd.y = __uint_as_float(__float_as_uint(xv[1]) & 0xFFFFE000u);
        hd.z = __uint_as_float(__float_as_uint(xv[2]) & 0xFFFFE000u);
        hd.w = __uint_as_float(__float_as_uint(xv[3]) & 0xFFFFE000u);
        lo_ = make_float4(xv[0] - hd.x, xv[1] - hd.y, xv[2] - hd.z, xv[3] - hd.w);
        *reinterpret_cast<float4*>(sm.A + tc_chunk_off(128, tid, c)) = hd;
        *reinterpret_cast<float4*>(sm.A + tc_chunk_off(128, tid, 4 + c)) = lo_;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    float dxa[4] = {0.f, 0.f, 0.f, 0.f};
    const int q = warp >> 2;                                   // column quarter of every accumulator / dims 4 q .. 4 q + 3 of Q
    const int sidx = 32 * (warp & 3) + lane;                   // state slot of this epilogue thread
    if (tid >= kBtEpi + 32) {
      // =============== bulk-copy producer (one thread): keeps both rings full, through this evaluation and into the next ===============
      if (lane == 0) {
        const long end_th = min(sm.total, b0 + n + kBtThStages), end_p = min(sm.total, b0 + n + kBtPStages);
        long fth = sm.ring[0], fp = sm.ring[1];
        unsigned long long idle_t0 = 0;
        int spins = 0;
        while (fth < end_th || fp < end_p) {
          bool any = false;
          if (fth < end_th && mbar_test(th_empty(sm, static_cast<int>(fth % kBtThStages)), static_cast<uint32_t>(((fth - kBtThStages) / kBtThStages) & 1))) {
            fetch_th(sm, g, fth++);
            any = true;
          }
          if (fp < end_p && mbar_test(p_empty(sm, static_cast<int>(fp % kBtPStages)), static_cast<uint32_t>(((fp - kBtPStages) / kBtPStages) & 1))) {
            fetch_p(sm, g, fp++);
            any = true;
          }
          if (any) spins = 0;
          else if (++spins > 64) {   // time-bounded like every other wait (common.cuh)
            if (spins == 65) idle_t0 = global_ns();
            __nanosleep(64);
            if (GPODE_WAIT_TIMEOUT_NS != 0ull && global_ns() - idle_t0 > GPODE_WAIT_TIMEOUT_NS) __trap();
          }
        }
        sm.ring[0] = fth;
        sm.ring[1] = fp;
      }
    } else if (tid >= kBtEpi) {
      // =============== MMA issuer (one thread) ===============
      if (lane == 0) {
#ifdef GPODE_BT_PROFILE
        long long pt[6] = {0, 0, 0, 0, 0, 0}, pc;
#define BT_T0 pc = clock64();
#define BT_T(i) { const long long now_ = clock64(); pt[i] += now_ - pc; pc = now_; }
#else
#define BT_T0
#define BT_T(i)
#endif
        BT_T0
        for (int i = 0; i < 3 && i < n; ++i) issue_theta(sm, g, b0 + i);
        BT_T(4)
        constexpr uint32_t idq = tc_idesc_bf16(128, kTcbQN);
        for (int i = 0; i < n; ++i) {
          const long b = b0 + i;
          const int acc = static_cast<int>(b % 3), ps = static_cast<int>(b % kBtPStages);
          const int j = i % nbi;
          const long kk = kk0 + i / nbi;
          tc_wait(tau_ready(sm, acc), static_cast<uint32_t>((b / 3) & 1));
          BT_T(0)
          if (j == 0 && kk >= 2) tc_wait(q_empty(sm, static_cast<int>(kk & 1)), static_cast<uint32_t>(((kk - 2) >> 1) & 1));
          BT_T(1)
          tc_wait(p_full(sm, ps), static_cast<uint32_t>((b / kBtPStages) & 1));
          BT_T(2)
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t dq = sm.tmem + kBtQCol + static_cast<uint32_t>(kk & 1) * 64, at = sm.tmem + acc * kTcbUnits;
          const uint32_t bp = smem_u32(sm.Bp + ps * kTcbPFloats);
#pragma unroll
          for (int ks = 0; ks < 2 * kTcbUnits / 16; ++ks)
            tc_mma_ts_bf16(dq, at + ks * 8, tc_desc(bp + ks * 2 * kTcbQN * 16, kTcbQN * 16), idq, (j > 0 || ks > 0) ? 1u : 0u);
          tc_commit(p_empty(sm, ps));
          if (j == nbi - 1) tc_commit(q_full(sm, static_cast<int>(kk & 1)));
          BT_T(3)
          if (i + 3 < n) issue_theta(sm, g, b + 3);
          BT_T(4)
        }
#ifdef GPODE_BT_PROFILE
        if (blockIdx.x == 5 && blockIdx.y == 0 && b0 == static_cast<long>(n))
          printf("bwd tc issuer (cycles per item over %d items): wait tau %lld, wait q_empty %lld, wait p tile %lld, issue Q %lld, issue theta (+ wait tile) %lld, refill %lld\n", n,
                 pt[0] / n, pt[1] / n, pt[2] / n, pt[3] / n, pt[4] / n, pt[5] / n);
#endif
      }
    } else {
      // =============== epilogue warps ===============
      const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
      (void)st;                                                // (st belongs to the state threads; every epilogue warp derives its own)
      const long n_state = static_cast<long>(blockIdx.x) * kBtStates + sidx;
      const bool live = n_state < g.N;
      const long s_glob = static_cast<long>(blockIdx.y) * g.N + (live ? n_state : g.N - 1);
      float Ak = 0.f, gk_cur = 0.f, gk_prev = 0.f, fv_cur = 0.f, fv_prev = 0.f;
      auto q_epilogue = [&](int k, float gk, float fv) {
        const long kk = kk0 + k;
        const int qb = static_cast<int>(kk & 1);
        tc_wait(q_full(sm, qb), static_cast<uint32_t>((kk >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tq = sm.tmem + kBtQCol + qb * 64 + lane_base;
        float qh[4], ql[4];
        tc_ld4(tq + 4 * q, qh);
        tc_ld4(tq + 24 + 4 * q, ql);
        const float es = tc_ld1(tq + 16) + tc_ld1(tq + 40);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc_arrive(q_empty(sm, qb));
        const float* hdr_k = sm.hdr + k * g.hdr_floats;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int d = 4 * q + t;
          float stat = 0.f;
          if (d < DP) {
            const float xv = sm.xs[d * kBtStates + sidx];
            const float dxk = gk * fmaf(2.f * hdr_k[d] * xv, es, qh[t] + ql[t]);
            dxa[t] += dxk;
            stat = xv * dxk;
          }
          stat = warp_sum(stat);
          if (lane == 0 && d < DP) atomicAdd(&sm.dell[k * DP + d], stat);
        }
        if (q == 0) {
          const float v = warp_sum(fv);
          if (lane == 0) atomicAdd(&sm.dvar[k], v);
        }
      };
#ifdef GPODE_BT_PROFILE
      long long et[4] = {0, 0, 0, 0}, ec = clock64();
#define BE_T(i) { const long long now_ = clock64(); et[i] += now_ - ec; ec = now_; }
#else
#define BE_T(i)
#endif
      for (int i = 0; i < n; ++i) {
        const long b = b0 + i;
        const int acc = static_cast<int>(b % 3);
        const int k = i / nbi, j = i - k * nbi;
        const bool is_k = j >= nbs;
        BE_T(3)
        if (j == 0) {
          const float* hdr_k = sm.hdr + k * g.hdr_floats;
          Ak = 0.f;
#pragma unroll
          for (int d = 0; d < DP; ++d) {
            const float xv = sm.xs[d * kBtStates + sidx];
            Ak = fmaf(hdr_k[d] * xv, xv, Ak);
          }
          gk_prev = gk_cur;
          fv_prev = fv_cur;
          const long at = k * kstride + s_glob * sstride;
          gk_cur = live ? gvec[at] : 0.f;
          fv_cur = (q == 0 && live) ? gk_cur * (fvec[at] - 0.5f * fpvec[at]) : 0.f;
        }
        tc_wait(acc_full(sm, acc), static_cast<uint32_t>((b / 3) & 1));
        BE_T(0)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ta = sm.tmem + acc * kTcbUnits + q * 32 + lane_base;
        uint32_t r0[16], r1[16];
        tc_ld16_async(ta, r0);
        tc_ld16_async(ta + 16, r1);
        tc_ld_wait(r0);
        if (is_k) {
#pragma unroll
          for (int v = 0; v < 16; ++v) r0[v] = tc_split_bf16(ex2_approx(__uint_as_float(r0[v]) + Ak));
        } else {
#pragma unroll
          for (int v = 0; v < 16; ++v) r0[v] = tc_split_bf16(__cosf(__uint_as_float(r0[v])));
        }
        tc_st16_async(ta, r0);
        tc_ld_wait(r1);
        if (is_k) {
#pragma unroll
          for (int v = 0; v < 16; ++v) r1[v] = tc_split_bf16(ex2_approx(__uint_as_float(r1[v]) + Ak));
        } else {
#pragma unroll
          for (int v = 0; v < 16; ++v) r1[v] = tc_split_bf16(__cosf(__uint_as_float(r1[v])));
        }
        tc_st16_async(ta + 16, r1);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc_arrive(tau_ready(sm, acc));
        BE_T(1)
        if (j == 0 && k > 0) q_epilogue(k - 1, gk_prev, fv_prev);   // deferred by one item: Q(k - 1) executes under this item's transcendentals
        BE_T(2)
      }
      q_epilogue(g.D_out - 1, gk_cur, fv_cur);
#ifdef GPODE_BT_PROFILE
      if (blockIdx.x == 5 && blockIdx.y == 0 && b0 == static_cast<long>(n) && (tid == 0 || tid == 480))
        printf("bwd tc epilogue warp %d (cycles per item): wait theta %lld, tau %lld, Q epilogue %lld, other %lld\n", warp, et[0] / n, et[1] / n, et[2] / n, et[3] / n);
#endif
    }
    sm.blk = b0 + n;
    sm.kk = kk0 + g.D_out;
    __syncthreads();   // every epilogue of this evaluation is done: xs may change, dx is complete in registers
    if (tid < kBtEpi) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (4 * q + t < DP) sm.dx[(4 * q + t) * kBtStates + sidx] = dxa[t];
    }
    __syncthreads();
  }
};

}  // namespace gpode
