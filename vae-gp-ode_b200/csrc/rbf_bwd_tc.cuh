// RBF reverse sweep FUSED with the parameter-gradient statistics on the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor
// memory), sm_100a, D > 8.  Structure of a fused attention backward (S = Q K^T -> P -> dQ = P K and dK = P^T Q from the SAME P):
//
// A CTA owns 128 states.  Per evaluation and output k the parameter rows arrive as operand tiles of 128 units ("items", k_rbf_pack_tcb:
// layout in rbf.h) in two halves through two bulk-copy rings.  Per item:
//   1. theta (128 states x 128 units, fp32 in tensor memory) = s (x . G + off): three kind::f16 k-steps (fp16 head / remainder split of
//      both operands, K = 16 = the input dimensions: X_h G_h + X_l G_h + X_h G_l, >= 21 bits) + one kind::tf32 k-step for the offsets
//      (K = 8: (s_n, s_n, 0 ..) x (off_h, off_l, 0 ..)).  s = s_n s_k are exact power-of-two block scales (per state from max |x_d|, per
//      output from max |coefficient|; both 1 for every shape of the reference) that keep the fp16 operands in range.
//   2. 16 epilogue warps (warp w: states 32 (w & 3).., units 32 (w >> 2)..) read theta (tcgen05.ld), evaluate
//          tau = -sin(theta)  (feature units; the + pi/2 is in the packed offset)      tau = 2^(theta + A_k(x))  (inducing units)
//      (un-scaling and A_k(x) are ONE FFMA), split tau into a bf16 head and a bf16 remainder (16 bits, fp32 exponent range) and store
//      both planes into a SHARED-memory tile laid out so that it is at once the K-major A operand of product 3 (rows = states) and the
//      MN-major A operand of product 4 (rows = units): 16-byte chunk (state s, plane p, units 8 c .. 8 c + 7) at (16 p + c) * 2048 + 16 s.
//   3. Q_k (128 states x 48) += tau B: 16 k-steps of kind::f16 (bf16), B = [P_h | P_l | w_h w_l ..] with P = weight x coefficient --
//      columns d and 16 + d sum to J^T contributions, 32 + 33 to the weighted sum of tau that the A_k(x) term needs.
//   4. inducing items only: PG (128 units x 48) = tau^T X': 16 k-steps over the CTA's states, X' = [g x_d heads | remainders | g_h g_l ..]
//      (one operand per output k, written by the drain warps, MN-major) -- columns d and 16 + d sum to sum_n g tau x_d, 32 + 33 to
//      sum_n g tau: exactly the accumulators the separate parameter-gradient pass (k_rbf_pgrad_mma: theta and the exponentials a third
//      time) used to fill.  The drain warps read PG as soon as it has executed and add it to the global accumulators (red.global.add.f32).
//   5. once per k the drain warps read Q_k: dx_k = g_k (Q + 2 c_d x_d Es), dx += dx_k, lengthscale statistic sum_n x_d dx_kd.
// Warp roles (24 warps): 16 epilogue warps do step 2 and nothing else (they pace the kernel); three MMA-issuer warps -- Q, PG, theta --
// each issue their products through one elected lane with provably warp-uniform control flow, so that descriptors live in uniform
// registers (a divergent `if (lane == 0)` makes ptxas wrap each tcgen05.mma in a ~100-cycle uniformisation loop -- measured,
// tools/tc_probe3.cu) and each block on their OWN hand-overs only (one issuer for all three spent a third of its time in such waits); one
// warp issues the bulk copies; four drain warps (one per tensor-memory lane quarter) write X' and do steps 4 (flush) and 5.  Flow
// control is mbarrier + tcgen05.commit, all waits time-bounded.
// Measured costs per item (tools/tc_probe3.cu): theta ~400 cycles, Q ~710, PG ~810 -- the second products are bound by reading the
// 64 KB tau tile from shared memory (128 B / cycle), not by the tensor pipe; the transcendentals need 1,024 MUFU cycles per item.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "rbf_kernels.cuh"
#include "tc_common.cuh"

namespace gpode {

#ifndef GPODE_BT_SLEEP_NS
#define GPODE_BT_SLEEP_NS 64   // poll interval of the epilogue warps' "I am ahead" waits (tau_empty, pg_full, q_full, xp_empty)
#endif
#ifndef GPODE_BT_A_TMEM
#define GPODE_BT_A_TMEM 1   // 1: the state operand of theta (X_h | X_l | block scales) is read from tensor memory instead of shared memory:
                            //    the kernel is bound by the L1 / shared-memory data pipe (97.6 % busy), and this operand, constant through an
                            //    evaluation, was re-read from shared memory by all four theta MMAs of every item (16 KB per item; 31.5 -> 31.0 ms)
#endif
constexpr int kBtACol = 448;                   // tensor-memory columns of that operand: 8 (X_h) + 8 (X_l) + 8 (scales)
#ifndef GPODE_BT_EXP
#define GPODE_BT_EXP 0   // timing experiments only (wrong results): 1 no PG MMAs, 2 no Q MMAs, 4 no transcendentals, 8 no tau stores, 16 no theta MMAs, 32 no proxy fence after the tau stores
#endif

constexpr int kBtStates = 128;                 // states per CTA
constexpr int kBtEpiWarps = 16;
constexpr int kBtEpi = kBtEpiWarps * 32;
constexpr int kBtFlushWarps = 4;               // drain warps: one per tensor-memory lane quarter (PG / Q accumulators -> global / shared memory)
constexpr int kBtThreads = kBtEpi + 64 + 32 * kBtFlushWarps + 64;   // + MMA-issuer warp (Q) + bulk-copy producer warp + drain warps + two more MMA-issuer warps (PG, theta)
constexpr int kBtPgIssuer = kBtEpiWarps + 2 + kBtFlushWarps;        // warp index of the PG issuer
constexpr int kBtThIssuer = kBtPgIssuer + 1;                        // ... of the theta issuer
constexpr int kBtTauBytes = 2 * 16 * 2048;     // one tau tile: 2 planes x 16 unit groups x (128 states x 16 B)
constexpr int kBtABytes = 3 * 4096;            // state operand of theta: X_h | X_l | (s_n, s_n, 0 ..)
constexpr int kBtXpBytes = 6 * 2048;           // X' (128 states x 48 columns bf16, MN-major: chunk (s, n / 8) at (n / 8) * 2048 + 16 s)
constexpr int kBtQCol = 256, kBtPgCol = 352;   // tensor-memory columns: 2 theta accumulators of 128, 2 Q of 48, 2 PG of 48
constexpr int kBtNBars = 26;
// byte offsets of the operand regions from the CTA's 1 KB-aligned shared-memory base (BwdTcSmem::sb)
constexpr uint32_t kBtOffTau = 0, kBtOffTh = 2 * kBtTauBytes, kBtOffP = kBtOffTh + 2 * kTcbThBytes, kBtOffA = kBtOffP + 2 * kTcbPBytes, kBtOffXp = kBtOffA + kBtABytes,
                   kBtOffXs = kBtOffXp + kBtXpBytes;

__host__ __device__ constexpr uint32_t tc_idesc_f16(int M, int N) {   // f32 accumulate, fp16 x fp16, both K-major
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t tc_idesc_bf16(int M, int N, int a_mn, int b_mn) {   // f32 accumulate, bf16 x bf16
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
// no-swizzle descriptor: lbo = byte step along K between core matrices, sbo = byte step along M / N between core matrices
__device__ __forceinline__ uint64_t tc_desc2(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((saddr & 0x3FFFF) >> 4) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) | (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (static_cast<uint64_t>(1) << 46);
}
// MMA / commit issued by ONE elected lane of a converged warp (operands warp-uniform)
__device__ __forceinline__ void tcu_mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p, e;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b),
               "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tcu_mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p, e;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b),
               "r"(idesc), "r"(acc)
               : "memory");
}
// A operand from tensor memory (rows = lanes), B from shared memory
__device__ __forceinline__ void tcu_mma_f16_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p, e;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem),
               "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tcu_mma_tf32_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p, e;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem),
               "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]),
               "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tcu_commit(uint32_t bar) {
  asm volatile("{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tc_ld2(uint32_t taddr, float (&v)[2]) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r0), "+r"(r1)::"memory");
  v[0] = __uint_as_float(r0);
  v[1] = __uint_as_float(r1);
}
__device__ __forceinline__ void tc_ld4_async(uint32_t taddr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
}
__device__ __forceinline__ void tc_ld2_async(uint32_t taddr, uint32_t& r0, uint32_t& r1) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
}
// orders later uses of asynchronously loaded registers behind the preceding tcgen05.wait::ld (volatile asm statements keep their order)
__device__ __forceinline__ void tc_ld_fence(uint32_t (&r)[10]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9])::"memory");
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
#if GPODE_BT_EXP
  if (GPODE_BT_EXP & 8) {
    asm volatile("" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d));
    return;
  }
#endif
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// (v0, v1) -> packed bf16 heads by truncation (element 0 in the low half) and packed bf16 remainders (rounded): |v - (h + l)| <= 2^-17 |v|
__device__ __forceinline__ void bt_split2(float v0, float v1, uint32_t& hd, uint32_t& rm) {
  hd = __byte_perm(__float_as_uint(v0), __float_as_uint(v1), 0x7632);
  // both remainders in ONE packed subtraction (FFMA2): the epilogue warps are issue-bound, every instruction saved per pair counts
  const float2 l = fma2(make_float2(__uint_as_float(__float_as_uint(v0) & 0xFFFF0000u), __uint_as_float(__float_as_uint(v1) & 0xFFFF0000u)), bc(-1.f), make_float2(v0, v1));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(rm) : "f"(l.y), "f"(l.x));
}

#ifdef GPODE_BT_PROFILE
#define BT_E0 long long et_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, ec_ = clock64();
#define BT_E(i) { const long long now_ = clock64(); et_[i] += now_ - ec_; ec_ = now_; }
#define BT_EPRINT if (blockIdx.x == 5 && blockIdx.y == 0 && b0 == static_cast<long>(n) && (tid == 0 || tid == 480)) \
  printf("bwd tc epilogue warp %d (cycles per item): first wait %lld, wait tau_empty %lld, ld issue + ld wait + arrives %lld, math+store %lld (+ wait next theta %lld), k setup + xprime %lld (+ wait xp_empty %lld), pg red %lld (+ waits pg_full / q_full %lld), q epilogue %lld (+ %lld)\n", warp, \
         et_[0] / n, et_[1] / n, et_[2] / n, et_[3] / n, et_[7] / n, et_[4] / n, et_[8] / n, et_[5] / n, et_[9] / n, et_[6] / n, et_[10] / n);
#define BT_I0 long long it_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ic_ = clock64();
#define BT_I(i) { const long long now_ = clock64(); it_[i] += now_ - ic_; ic_ = now_; }
#define BT_IPRINT if (blockIdx.x == 5 && blockIdx.y == 0 && b0 == static_cast<long>(n) && lane == 0) \
  printf("bwd tc issuer (cycles per item): theta waits (tile, acc_empty) %lld, theta issue %lld, wait tau_full %lld, wait p tile %lld, wait q_empty/xp/pg_empty %lld, issue Q %lld, issue PG %lld\n", \
         it_[0] / n, it_[1] / n, it_[2] / n, it_[3] / n, it_[4] / n, it_[5] / n, it_[6] / n);
#else
#define BT_E0
#define BT_E(i)
#define BT_EPRINT
#define BT_I0
#define BT_I(i)
#define BT_IPRINT
#endif

struct BwdTcSmem {
  unsigned char* tau;   // 2 x kBtTauBytes
  unsigned char* th;    // 2 x kTcbThBytes   theta-half ring
  unsigned char* pp;    // 2 x kTcbPBytes    second-half ring
  unsigned char* A;     // kBtABytes
  unsigned char* xp;    // kBtXpBytes; after the last MMA of an evaluation its first 8 KB double as dx (J^T g, [DP][128] floats)
  float* xs;            // [16][128] staged states (generic solver glue), stride kBtStates
  float* dx;            // = xp
  float* hdr;           // [D_out][hdr_floats]
  float* dell;          // [D_out][DP] lengthscale statistic, then dvar [D_out] (contiguous, like SweepSmem)
  float* dvar;
  float* sk;            // [D_out] 1 / s_k
  float* gs;            // [16][128] upstream gradient g_k of the CTA's states for the evaluation in flight (staged by the state threads)
  float* g_pg;          // global accumulators of the launch (bind()): sum_n g tau x_d and sum_n g tau per inducing unit
  float* g_dnu;
  uint64_t* bars;
  uint32_t sb;          // shared-window address of the region base (opaque to the compiler: every operand address is sb + constant)
  uint32_t sbars;       // ... of bars
  uint32_t* tmem_slot;
  const float* tiles;   // operand tiles of this sample (global)
  uint32_t tmem;
  long blk;             // running item counter of this CTA
  long kk;              // running (evaluation, k) counter: Q buffer kk & 1
  long pgc;             // running inducing-item counter: PG buffer pgc & 1
  long total;
};

inline int rbf_bwd_tc_smem_bytes(const RbfGeom& g) {
  return 2 * kBtTauBytes + 2 * kTcbThBytes + 2 * kTcbPBytes + kBtABytes + kBtXpBytes + 2 * 16 * kBtStates * 4 + (g.D_out * g.hdr_floats + g.D_out * (g.DP + 1) + g.D_out + 8) * 4 +
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
  static constexpr bool kCoopGlue = true;   // sweep.cuh: the solver glue between evaluations is spread over all 576 threads
  static constexpr int kThreadsBwd = kBtThreads;
  static constexpr int kMinBlocksBwd = 1;
  using Geom = RbfGeom;
  using Accum = RbfAccum;
  using Smem = BwdTcSmem;
  static_assert(DP_ <= 16, "one 16-wide K block per operand part");

  // ring / accumulator barriers (index = slot or buffer 0 / 1)
  __device__ static __forceinline__ uint32_t th_full(const Smem& sm, int s) { return sm.sbars + 8u * static_cast<uint32_t>(s); }
  __device__ static __forceinline__ uint32_t th_empty(const Smem& sm, int s) { return sm.sbars + 8u * static_cast<uint32_t>(2 + s); }
  __device__ static __forceinline__ uint32_t p_full(const Smem& sm, int s) { return sm.sbars + 8u * static_cast<uint32_t>(4 + s); }
  __device__ static __forceinline__ uint32_t p_empty(const Smem& sm, int s) { return sm.sbars + 8u * static_cast<uint32_t>(6 + s); }
  __device__ static __forceinline__ uint32_t acc_full(const Smem& sm, int a) { return sm.sbars + 8u * static_cast<uint32_t>(8 + a); }     // theta(b) executed
  __device__ static __forceinline__ uint32_t acc_empty(const Smem& sm, int a) { return sm.sbars + 8u * static_cast<uint32_t>(10 + a); }   // every warp has read theta(b)
  __device__ static __forceinline__ uint32_t tau_full(const Smem& sm, int a) { return sm.sbars + 8u * static_cast<uint32_t>(12 + a); }    // every warp has stored tau(b)
  __device__ static __forceinline__ uint32_t tau_empty(const Smem& sm, int a) { return sm.sbars + 8u * static_cast<uint32_t>(14 + a); }   // Q(b) and PG(b) executed
  __device__ static __forceinline__ uint32_t q_full(const Smem& sm, int q) { return sm.sbars + 8u * static_cast<uint32_t>(16 + q); }
  __device__ static __forceinline__ uint32_t q_empty(const Smem& sm, int q) { return sm.sbars + 8u * static_cast<uint32_t>(18 + q); }
  __device__ static __forceinline__ uint32_t pg_full(const Smem& sm, int q) { return sm.sbars + 8u * static_cast<uint32_t>(20 + q); }
  __device__ static __forceinline__ uint32_t pg_empty(const Smem& sm, int q) { return sm.sbars + 8u * static_cast<uint32_t>(22 + q); }
  __device__ static __forceinline__ uint32_t xp_full(const Smem& sm) { return sm.sbars + 8u * 24u; }                // X' of this k written
  __device__ static __forceinline__ uint32_t xp_empty(const Smem& sm) { return sm.sbars + 8u * 25u; }               // last PG of this k executed

  __device__ static __forceinline__ Smem carve(float* smem, const Geom& g) {
    Smem s;
    // (the 1 KB alignment is applied as an OFFSET to the shared-memory array: a pointer rebuilt from an integer loses its address space,
    //  and every access through it becomes a generic LD / ST with a window check -- measured: the epilogue loop was full of them)
    uint32_t pad = (1024u - (smem_u32(smem) & 1023u)) & 1023u;
    asm volatile("" : "+r"(pad));   // opaque: never re-derived
    unsigned char* base = reinterpret_cast<unsigned char*>(smem) + pad;
    s.sb = smem_u32(base);
    asm volatile("" : "+r"(s.sb));
    s.tau = base;
    s.th = s.tau + 2 * kBtTauBytes;
    s.pp = s.th + 2 * kTcbThBytes;
    s.A = s.pp + 2 * kTcbPBytes;
    s.xp = s.A + kBtABytes;
    s.xs = reinterpret_cast<float*>(s.xp + kBtXpBytes);
    s.dx = reinterpret_cast<float*>(s.xp);
    s.hdr = s.xs + 16 * kBtStates;
    s.dell = s.hdr + g.D_out * g.hdr_floats;
    s.dvar = s.dell + g.D_out * DP;
    s.sk = s.dvar + g.D_out;
    s.gs = s.sk + ((g.D_out + 3) & ~3);
    float* end = s.gs + 16 * kBtStates;   // (base is 1 KB aligned and every block before this is a multiple of 8 bytes)
    s.bars = reinterpret_cast<uint64_t*>(end + ((g.D_out * g.hdr_floats + g.D_out * (DP + 1)) & 1));
    s.sbars = smem_u32(s.bars);
    asm volatile("" : "+r"(s.sbars));
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.bars + kBtNBars);
    s.g_pg = nullptr;
    s.g_dnu = nullptr;
    s.blk = 0;
    s.kk = 0;
    s.pgc = 0;
    return s;
  }

  __device__ static __forceinline__ void fetch_th(const Smem& sm, const Geom& g, long b) {
    const int per_eval = g.D_out * rbf_tcb_items(g);
    const int slot = static_cast<int>(b & 1);
    const float* src = sm.tiles + static_cast<size_t>(b % per_eval) * kTcbTileFloats;
    mbar_expect_tx(th_full(sm, slot), kTcbThBytes);
    bulk_g2s(sm.sb + kBtOffTh + slot * kTcbThBytes, src, kTcbThBytes, th_full(sm, slot));
  }
  __device__ static __forceinline__ void fetch_p(const Smem& sm, const Geom& g, long b) {
    const int per_eval = g.D_out * rbf_tcb_items(g);
    const int slot = static_cast<int>(b & 1);
    const float* src = sm.tiles + static_cast<size_t>(b % per_eval) * kTcbTileFloats + kTcbThFloats;
    mbar_expect_tx(p_full(sm, slot), kTcbPBytes);
    bulk_g2s(sm.sb + kBtOffP + slot * kTcbPBytes, src, kTcbPBytes, p_full(sm, slot));
  }

  __device__ static __forceinline__ long setup(Smem& sm, ChunkPipe&, const Geom& g, const float* packed, long n_evals, bool) {
    const int l = blockIdx.y, tid = threadIdx.x;
    const float* hdr = rbf_hdr_ptr(packed, g, l);
    for (int i = tid; i < g.D_out * g.hdr_floats; i += blockDim.x) sm.hdr[i] = hdr[i];
    for (int i = tid; i < 16 * kBtStates; i += blockDim.x) sm.xs[i] = 0.f;
    for (int i = tid; i < kBtXpBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm.xp)[i] = 0u;
    for (int i = tid; i < kBtABytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm.A)[i] = 0u;
    for (int i = tid; i < g.D_out * (DP + 1); i += blockDim.x) sm.dell[i] = 0.f;  // dell and dvar
    for (int i = tid; i < g.D_out; i += blockDim.x) {
      float s, inv;
      rbf_pow2_scale(rbf_maxabs_ptr(packed, g, l)[i], s, inv);
      sm.sk[i] = inv;
    }
    sm.tiles = rbf_tcb_tiles_ptr(packed, g, l);
    sm.total = n_evals * g.D_out * rbf_tcb_items(g);
    if (tid == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(th_full(sm, i), 1);
        mbar_init(th_empty(sm, i), 1);
        mbar_init(p_full(sm, i), 1);
        mbar_init(p_empty(sm, i), 1);
        mbar_init(acc_full(sm, i), 1);
        mbar_init(acc_empty(sm, i), kBtEpiWarps);
        mbar_init(tau_full(sm, i), kBtEpiWarps);
        mbar_init(tau_empty(sm, i), 2);   // both issuer warps: Q(b) executed, PG(b) executed (or nothing to do)
        mbar_init(q_full(sm, i), 1);
        mbar_init(q_empty(sm, i), kBtFlushWarps);
        mbar_init(pg_full(sm, i), 1);
        mbar_init(pg_empty(sm, i), kBtFlushWarps);
      }
      mbar_init(xp_full(sm), kBtFlushWarps);
      mbar_init(xp_empty(sm), 1);
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
    if (tid == kBtEpi + 32) {   // the producer thread owns the rings: first tiles of both
      for (long f = 0; f < 2 && f < sm.total; ++f) {
        fetch_th(sm, g, f);
        fetch_p(sm, g, f);
      }
    }
    return sm.total;
  }

  __device__ static __forceinline__ void finish(Smem&) {}

  __device__ static __forceinline__ void flush(const Smem& sm, const Geom& g, const Accum& acc) {
    // (the sweep kernels synchronise the CTA before flush(): the shared-memory accumulators are complete)
    for (int i = threadIdx.x; i < g.D_out * DP; i += blockDim.x) atomicAdd(&acc.dell_x[i], sm.dell[i]);
    for (int i = threadIdx.x; i < g.D_out; i += blockDim.x) atomicAdd(&acc.dvar[i], sm.dvar[i]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sm.tmem), "r"(512) : "memory");
  }

  // the sweep kernels hand the launch's accumulators to policies that fill them during the sweep (sweep.cuh: bind_accum)
  __device__ static __forceinline__ void bind(Smem& sm, const Accum& acc) {
    sm.g_pg = acc.pg;
    sm.g_dnu = acc.dnu;
  }

  __device__ static __forceinline__ void vjp(ChunkPipe&, const Geom& g, long, Smem& sm, const States<R>& st, const float* gvec, const float* fvec,
                                             const float* fpvec, long kstride, long sstride) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: the role branches may use the uniform datapath
    const int nbs = rbf_tcb_items_s(g), nbi = rbf_tcb_items(g), nbm = nbi - nbs;
    const int n = g.D_out * nbi;   // items of one evaluation
    const long b0 = sm.blk, kk0 = sm.kk, pg0 = sm.pgc;
    (void)st;
    if (tid < kBtStates) {   // ---- this state's operand row: block scale, fp16 heads and remainders, (s_n, s_n) for the offsets ----
      float xv[16], mx = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        xv[d] = d < DP ? sm.xs[d * kBtStates + tid] : 0.f;
        mx = fmaxf(mx, fabsf(xv[d]));
      }
      float sn, inv;
      rbf_pow2_scale(mx, sn, inv);
#if !GPODE_BT_A_TMEM
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t hh[4], ll[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float v0 = xv[8 * c + 2 * i] * sn, v1 = xv[8 * c + 2 * i + 1] * sn;
          const __half2 h = __floats2half2_rn(v0, v1);
          const __half2 lo_ = __floats2half2_rn(v0 - __low2float(h), v1 - __high2float(h));
          hh[i] = *reinterpret_cast<const uint32_t*>(&h);
          ll[i] = *reinterpret_cast<const uint32_t*>(&lo_);
        }
        sts128(sm.sb + kBtOffA + c * 2048 + tid * 16, hh[0], hh[1], hh[2], hh[3]);
        sts128(sm.sb + kBtOffA + 4096 + c * 2048 + tid * 16, ll[0], ll[1], ll[2], ll[3]);
      }
      sts128(sm.sb + kBtOffA + 8192 + tid * 16, __float_as_uint(sn), __float_as_uint(sn), 0u, 0u);
#endif
#if GPODE_BT_A_TMEM
      {   // this state's operand row into tensor memory (lane = state): fp16 pairs of X_h, of X_l, then the tf32 row (s_n, s_n, 0 ..)
        uint32_t th_[8], tl_[8], ts_[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float v0 = xv[2 * i] * sn, v1 = xv[2 * i + 1] * sn;
          const __half2 h = __floats2half2_rn(v0, v1);
          const __half2 lo_ = __floats2half2_rn(v0 - __low2float(h), v1 - __high2float(h));
          th_[i] = *reinterpret_cast<const uint32_t*>(&h);
          tl_[i] = *reinterpret_cast<const uint32_t*>(&lo_);
          ts_[i] = i < 2 ? __float_as_uint(sn) : 0u;
        }
        const uint32_t ta_ = sm.tmem + kBtACol + (static_cast<uint32_t>((tid >> 5) * 32) << 16);
        tc_st8(ta_, th_);
        tc_st8(ta_ + 8, tl_);
        tc_st8(ta_ + 16, ts_);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
#endif
    }
    if (tid < kBtEpi) {
      // the upstream gradient of this evaluation goes to shared memory here (the item loop below must not hold global loads in flight:
      // with 96 registers their results were spilled on arrival, i.e. the warp sat out the full DRAM latency once per output), and the
      // variance statistic sum_n g_k (f_k - f_p,k / 2) is finished on the spot.  Element (k, state slot) <-> thread: <= 4 independent
      // loads per thread, one round trip; a warp's lanes share k.
#pragma unroll 4
      for (int e = tid; e < g.D_out * kBtStates; e += kBtEpi) {
        const int k = e / kBtStates, sl = e - k * kBtStates;
        const long n_st = static_cast<long>(blockIdx.x) * kBtStates + sl;
        const bool live_st = n_st < g.N;
        const long at = k * kstride + (static_cast<long>(blockIdx.y) * g.N + (live_st ? n_st : g.N - 1)) * sstride;
        const float gk = live_st ? gvec[at] : 0.f;
        sm.gs[e] = gk;
        const float v = warp_sum(gk * (fvec[at] - 0.5f * fpvec[at]));
        if (lane == 0) atomicAdd(&sm.dvar[k], v);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int q = warp >> 2;                                   // unit quarter of every item / dims 4 q .. 4 q + 3 of Q and PG
    int sidx = 32 * (warp & 3) + lane;                         // state slot (Q) / unit slot (PG) of this epilogue thread
    asm volatile("" : "+r"(sidx));                             // (opaque: otherwise re-derived from S2R SR_TID.X inside the item loop)
    if (warp == kBtEpiWarps + 1) {
      // =============== bulk-copy producer: keeps both rings full, through this evaluation and into the next ===============
      if (lane == 0) {
        // The slots drain in the issuer's own order -- theta(b + 2) then Q(b) per item -- so the producer blocks on them in that order
        // (hardware wake-up, no polling): tile f + 2 of the theta ring follows theta(f), tile f of the second ring follows Q(f - 2).
        const long end = min(sm.total, b0 + n + 2);
        for (long f = b0; f < end; ++f) {   // tiles < b0 + 2 were fetched by setup() or by the tail of the previous evaluation
          if (f + 2 < end) {
            tc_wait(th_empty(sm, static_cast<int>(f & 1)), static_cast<uint32_t>((f >> 1) & 1));
            fetch_th(sm, g, f + 2);
          }
          if (f >= b0 + 2) {
            tc_wait(p_empty(sm, static_cast<int>(f & 1)), static_cast<uint32_t>(((f - 2) >> 1) & 1));
            fetch_p(sm, g, f);
          }
        }
      }
    } else if (warp == kBtEpiWarps) {
      // =============== MMA issuer: the whole warp walks the schedule, one elected lane issues ===============
      const uint32_t aTau = sm.sb + kBtOffTau, aP = sm.sb + kBtOffP;
      constexpr uint32_t id_q = tc_idesc_bf16(128, kTcbQN, 0, 0);
      BT_I0
      int k = 0, j = 0;
#pragma unroll 1
      for (int i = 0; i < n; ++i) {
        const long b = b0 + i;
        const int slot = static_cast<int>(b & 1);
        const long kk = kk0 + k;
        tc_wait(tau_full(sm, slot), static_cast<uint32_t>((b >> 1) & 1));
        BT_I(2)
        tc_wait(p_full(sm, slot), static_cast<uint32_t>((b >> 1) & 1));
        BT_I(3)
        if (j == 0 && kk >= 2) tc_wait(q_empty(sm, static_cast<int>(kk & 1)), static_cast<uint32_t>(((kk - 2) >> 1) & 1));
        BT_I(4)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ta = aTau + slot * kBtTauBytes, bp = aP + slot * kTcbPBytes;
        const uint32_t dq = sm.tmem + kBtQCol + static_cast<uint32_t>(kk & 1) * kTcbQN;
#pragma unroll
        for (int s = 0; s < ((GPODE_BT_EXP & 2) ? 0 : 16); ++s)   // k-step s: plane s / 8, units 16 (s % 8) ..
          tcu_mma_f16(dq, tc_desc2(ta + s * 2 * 2048, 2048, 128), tc_desc2(bp + (s & 7) * 2 * kTcbQN * 16, kTcbQN * 16, 128), id_q, (j > 0 || s > 0) ? 1u : 0u);
        tcu_commit(p_empty(sm, slot));
        if (j == nbi - 1) tcu_commit(q_full(sm, static_cast<int>(kk & 1)));
        BT_I(5)
        tcu_commit(tau_empty(sm, slot));
        BT_I(6)
        if (++j == nbi) {
          j = 0;
          ++k;
        }
      }
      BT_IPRINT
    } else if (warp == kBtThIssuer) {
      // =============== third MMA issuer: theta(b) = s (x . G + off), two items ahead of the epilogue warps ===============
      // theta(b + 2) goes out as soon as the epilogue warps have READ theta(b) (mid-item: its accumulator is free) and its operand tile has
      // landed; on a warp of its own these waits stall nothing else.
      const uint32_t aA = sm.sb + kBtOffA, aTh = sm.sb + kBtOffTh;
      constexpr uint32_t id_th = tc_idesc_f16(128, kTcbUnits), id_off = tc_idesc(128, kTcbUnits);
      BT_I0
      auto issue_theta = [&](long b) {
        const int slot = static_cast<int>(b & 1);
        tc_wait(th_full(sm, slot), static_cast<uint32_t>((b >> 1) & 1));
        if (b >= 2) tc_wait(acc_empty(sm, slot), static_cast<uint32_t>(((b - 2) >> 1) & 1));
        BT_I(0)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = sm.tmem + slot * kTcbUnits, bt = aTh + slot * kTcbThBytes;
        if (!(GPODE_BT_EXP & 16)) {
#if GPODE_BT_A_TMEM
        tcu_mma_f16_ts(d, sm.tmem + kBtACol, tc_desc2(bt, 2048, 128), id_th, 0u);                      // X_h G_h
        tcu_mma_f16_ts(d, sm.tmem + kBtACol + 8, tc_desc2(bt, 2048, 128), id_th, 1u);                  // X_l G_h
        tcu_mma_f16_ts(d, sm.tmem + kBtACol, tc_desc2(bt + 4096, 2048, 128), id_th, 1u);               // X_h G_l
        tcu_mma_tf32_ts(d, sm.tmem + kBtACol + 16, tc_desc2(bt + 8192, 2048, 128), id_off, 1u);        // (s_n, s_n) x (off_h, off_l)
        (void)aA;
#else
        tcu_mma_f16(d, tc_desc2(aA, 2048, 128), tc_desc2(bt, 2048, 128), id_th, 0u);                    // X_h G_h
        tcu_mma_f16(d, tc_desc2(aA + 4096, 2048, 128), tc_desc2(bt, 2048, 128), id_th, 1u);            // X_l G_h
        tcu_mma_f16(d, tc_desc2(aA, 2048, 128), tc_desc2(bt + 4096, 2048, 128), id_th, 1u);            // X_h G_l
        tcu_mma_tf32(d, tc_desc2(aA + 8192, 2048, 128), tc_desc2(bt + 8192, 2048, 128), id_off, 1u);   // (s_n, s_n) x (off_h, off_l)
#endif
        }
        tcu_commit(acc_full(sm, slot));
        tcu_commit(th_empty(sm, slot));
        BT_I(1)
      };
#pragma unroll 1
      for (int i = 0; i < n; ++i) issue_theta(b0 + i);
    } else if (warp == kBtPgIssuer) {
      // =============== second MMA issuer: the parameter-gradient product PG(b) = tau(b)^T X' of the inducing items ===============
      // Its dependencies (X' of the output, the PG accumulator two tiles back) are handed over by the drain warps; blocking on them in
      // the first issuer stalled theta and Q behind them (measured: a third of that warp's time was such waits).
      const uint32_t aTau = sm.sb + kBtOffTau, aXp = sm.sb + kBtOffXp;
      constexpr uint32_t id_pg = tc_idesc_bf16(128, kTcbQN, 1, 1);
      int k = 0, j = 0;
#pragma unroll 1
      for (int i = 0; i < n; ++i) {
        const long b = b0 + i;
        const int slot = static_cast<int>(b & 1);
        const long kk = kk0 + k;
        tc_wait(tau_full(sm, slot), static_cast<uint32_t>((b >> 1) & 1));
        if (j >= nbs) {
          const long pc = pg0 + static_cast<long>(k) * nbm + (j - nbs);
          if (j == nbs) tc_wait(xp_full(sm), static_cast<uint32_t>(kk & 1));
          if (pc >= 2) tc_wait(pg_empty(sm, static_cast<int>(pc & 1)), static_cast<uint32_t>(((pc - 2) >> 1) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ta = aTau + slot * kBtTauBytes;
          const uint32_t dpg = sm.tmem + kBtPgCol + static_cast<uint32_t>(pc & 1) * kTcbQN;
#pragma unroll
          for (int p = 0; p < ((GPODE_BT_EXP & 1) ? 0 : 2); ++p)
#pragma unroll
            for (int s = 0; s < 8; ++s)   // k-step: states 16 s .. 16 s + 15; A = tau^T (MN-major: next 8 states + 128 B, next 8 units + 2048 B)
              tcu_mma_f16(dpg, tc_desc2(ta + p * 16 * 2048 + s * 2 * 128, 128, 2048), tc_desc2(aXp + s * 2 * 128, 128, 2048), id_pg, (p > 0 || s > 0) ? 1u : 0u);
          tcu_commit(pg_full(sm, static_cast<int>(pc & 1)));
          if (j == nbi - 1) tcu_commit(xp_empty(sm));
        }
        tcu_commit(tau_empty(sm, slot));   // (feature items: no MMA of this warp is pending, the arrival is immediate)
        if (++j == nbi) {
          j = 0;
          ++k;
        }
      }
    } else if (warp >= kBtEpiWarps + 2 && warp < kBtPgIssuer) {
      // =============== drain warps (one per tensor-memory lane quarter; thread <-> unit of a PG tile / state of a Q tile) ===============
      // Everything that is not theta -> tau runs here, off the 16 epilogue warps that pace the kernel: X' of every output (the B operand of
      // the parameter-gradient product), the PG tiles (tensor memory -> red.global into the launch's accumulators) and, once per output,
      // the state gradient with its lengthscale statistic.  (Measured before the split: a fifth of the epilogue warps' time.)
      const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
      const uint32_t aXp = sm.sb + kBtOffXp;
      const uint32_t tq0 = sm.tmem + kBtQCol + lane_base, tp0 = sm.tmem + kBtPgCol + lane_base;
      float dxa[16];
#pragma unroll
      for (int d = 0; d < 16; ++d) dxa[d] = 0.f;
      // X' of output k (MN-major B operand of PG): columns d (heads of g x_d), 16 + d (remainders), 32 / 33 (g); the last PG of the
      // previous output must have executed.  (x is re-read from shared memory where it is needed: registers are short in these warps.)
      auto write_xprime = [&](int k, long kk) {
        if (kk >= 1) mbar_wait_sleepy(xp_empty(sm), static_cast<uint32_t>((kk - 1) & 1), GPODE_BT_SLEEP_NS);
        const float gk = sm.gs[k * kBtStates + sidx];
        const uint32_t row = aXp + sidx * 16;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t hd[4], rm[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int d = 8 * c + 2 * v;
            bt_split2(d < DP ? gk * sm.xs[d * kBtStates + sidx] : 0.f, d + 1 < DP ? gk * sm.xs[(d + 1) * kBtStates + sidx] : 0.f, hd[v], rm[v]);
          }
          sts128(row + c * 2048, hd[0], hd[1], hd[2], hd[3]);
          sts128(row + (2 + c) * 2048, rm[0], rm[1], rm[2], rm[3]);
        }
        uint32_t hg, lg;
        bt_split2(gk, 0.f, hg, lg);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(row + 4 * 2048), "r"((hg & 0xFFFFu) | (lg << 16)) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc_arrive(xp_full(sm));
      };
      write_xprime(0, kk0);
      long pc = pg0;
#pragma unroll 1
      for (int k = 0; k < g.D_out; ++k) {
        const long kk = kk0 + k;
        // ---- the PG tiles of this k: tensor memory -> global accumulators ----
#pragma unroll 1
        for (int jm = 0; jm < nbm; ++jm, ++pc) {
          mbar_wait_sleepy(pg_full(sm, static_cast<int>(pc & 1)), static_cast<uint32_t>((pc >> 1) & 1), GPODE_BT_SLEEP_NS);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tp = tp0 + static_cast<uint32_t>(pc & 1) * kTcbQN;
          uint32_t ph[16], pl[16], ps0, ps1;
          tc_ld16_async(tp, ph);
          tc_ld16_async(tp + 16, pl);
          tc_ld2_async(tp + 32, ps0, ps1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          asm volatile("" : "+r"(ps0), "+r"(ps1)::"memory");
          tc_ld_wait(ph);   // (orders the uses below behind the wait)
          tc_ld_wait(pl);
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) tc_arrive(pg_empty(sm, static_cast<int>(pc & 1)));
          const int unit = jm * kTcbUnits + sidx;
          if (unit < g.M) {
            const size_t base = (static_cast<size_t>(blockIdx.y) * g.D_out + k) * (2 * g.MP2) + unit;
            float* dst = sm.g_pg + base * DP;
            if constexpr (DP == 16) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 4 * c), "f"(__uint_as_float(ph[4 * c]) + __uint_as_float(pl[4 * c])),
                             "f"(__uint_as_float(ph[4 * c + 1]) + __uint_as_float(pl[4 * c + 1])), "f"(__uint_as_float(ph[4 * c + 2]) + __uint_as_float(pl[4 * c + 2])),
                             "f"(__uint_as_float(ph[4 * c + 3]) + __uint_as_float(pl[4 * c + 3]))
                             : "memory");
            } else {
#pragma unroll
              for (int d = 0; d < DP; ++d) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + d), "f"(__uint_as_float(ph[d]) + __uint_as_float(pl[d])) : "memory");
            }
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(sm.g_dnu + base), "f"(__uint_as_float(ps0) + __uint_as_float(ps1)) : "memory");
          }
        }
        // ---- X' of the NEXT output first (the issuer needs it two items into that output), then Q of this k: state gradient
        //      dx_k = g_k (Q + 2 c_d x_d Es) and the lengthscale statistic sum_n x_d dx_kd ----
        if (k + 1 < g.D_out) write_xprime(k + 1, kk + 1);
        const float gk = sm.gs[k * kBtStates + sidx];
        mbar_wait_sleepy(q_full(sm, static_cast<int>(kk & 1)), static_cast<uint32_t>((kk >> 1) & 1), GPODE_BT_SLEEP_NS);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
          const uint32_t tq = tq0 + static_cast<uint32_t>(kk & 1) * kTcbQN;
          uint32_t qh[16], ql[16], qs0, qs1;
          tc_ld16_async(tq, qh);
          tc_ld16_async(tq + 16, ql);
          tc_ld2_async(tq + 32, qs0, qs1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          asm volatile("" : "+r"(qs0), "+r"(qs1)::"memory");
          tc_ld_wait(qh);
          tc_ld_wait(ql);
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) tc_arrive(q_empty(sm, static_cast<int>(kk & 1)));
          const float es = __uint_as_float(qs0) + __uint_as_float(qs1);
          const float* hdr_k = sm.hdr + k * g.hdr_floats;
          float red[16];
#pragma unroll
          for (int d = 0; d < 16; ++d) {
            red[d] = 0.f;
            if (d < DP) {
              const float xd = sm.xs[d * kBtStates + sidx];
              const float dxk = gk * fmaf(2.f * hdr_k[d] * xd, es, __uint_as_float(qh[d]) + __uint_as_float(ql[d]));
              dxa[d] += dxk;
              red[d] = xd * dxk;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)   // DP independent butterflies, interleaved
#pragma unroll
            for (int d = 0; d < DP; ++d) red[d] += __shfl_xor_sync(0xffffffffu, red[d], o);
          float rv = 0.f;
#pragma unroll
          for (int d = 0; d < DP; ++d) rv = lane == d ? red[d] : rv;
          if (lane < DP) atomicAdd(&sm.dell[k * DP + lane], rv);   // shared-memory accumulator of the CTA (flush() adds it to the launch's)
        }
      }
      sm.blk = b0 + n;
      sm.kk = kk0 + g.D_out;
      sm.pgc = pg0 + static_cast<long>(g.D_out) * nbm;
      __syncthreads();   // every MMA of this evaluation has executed and every epilogue is done: xs / X' may change (dx aliases X')
#pragma unroll
      for (int d = 0; d < DP; ++d) sm.dx[d * kBtStates + sidx] = dxa[d];
      __syncthreads();
      return;
    } else {
      // =============== epilogue warps: theta -> tau, nothing else ===============
      const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
      float mx = 0.f;
#pragma unroll
      for (int d = 0; d < DP; ++d) mx = fmaxf(mx, fabsf(sm.xs[d * kBtStates + sidx]));
      float sn_, inv_n;
      rbf_pow2_scale(mx, sn_, inv_n);   // this state's block scale (as the state threads computed it)
      float Ak = 0.f, inv_s = 1.f;
      const uint32_t aTau = sm.sb + kBtOffTau;
      const uint32_t ta0 = sm.tmem + q * 32 + lane_base;
      int k = 0, j = 0;              // coordinates of item i
      uint32_t rA[16], rB[16];
      const uint32_t b0u = static_cast<uint32_t>(b0);   // (slots and phase bits need the low bits only; 64-bit counters were spilled in this loop)
      int i_wait = b0 >= 2 ? 0 : static_cast<int>(2 - b0);   // first item of this evaluation whose tau buffer was used before
      asm volatile("" : "+r"(i_wait));                              // (opaque: not re-derived from the spilled 64-bit counter in the loop)
      tc_wait(acc_full(sm, static_cast<int>(b0u & 1)), (b0u >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      tc_ld16_async(ta0 + (b0u & 1) * kTcbUnits, rA);   // first half of the first item; later ones are prefetched below
      uint32_t b = b0u;              // running item index (low bits)
#pragma unroll 1
      for (int i = 0; i < n; ++i, ++b) {
        const int slot = static_cast<int>(b & 1);
        const bool is_k = j >= nbs;
        if (j == 0) {
          const float* hdr_k = sm.hdr + k * g.hdr_floats;
          Ak = 0.f;
#pragma unroll
          for (int d = 0; d < DP; ++d) {
            const float xv = sm.xs[d * kBtStates + sidx];
            Ak = fmaf(hdr_k[d] * xv, xv, Ak);
          }
          inv_s = inv_n * sm.sk[k];
        }
        // ---- the first theta half was prefetched under the previous item; the second half is loaded under the first half's
        //      transcendentals ----
        if (i >= i_wait) tc_wait(tau_empty(sm, slot), ((b - 2u) >> 1) & 1u);   // (on the critical path since the drain warps exist: hardware wake-up, no sleep)
        tc_ld_wait(rA);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tc_ld16_async(ta0 + slot * kTcbUnits + 16, rB);
        const float addk = is_k ? Ak : 0.f;
        const uint32_t trow = aTau + slot * kBtTauBytes + sidx * 16 + (4 * q) * 2048;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t hd[4], rm[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float2 tt = fma2(make_float2(__uint_as_float(rA[8 * c + 2 * v]), __uint_as_float(rA[8 * c + 2 * v + 1])), bc(inv_s), bc(addk));   // un-scaling and A_k(x): one packed FFMA2 per pair
            const float t0 = tt.x, t1 = tt.y;
            bt_split2((GPODE_BT_EXP & 4) ? t0 : (is_k ? ex2_approx(t0) : __cosf(t0)), (GPODE_BT_EXP & 4) ? t1 : (is_k ? ex2_approx(t1) : __cosf(t1)), hd[v], rm[v]);
          }
          sts128(trow + c * 2048, hd[0], hd[1], hd[2], hd[3]);
          sts128(trow + (16 + c) * 2048, rm[0], rm[1], rm[2], rm[3]);
        }
        tc_ld_wait(rB);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc_arrive(acc_empty(sm, slot));
        if (i + 1 < n) {   // first half of the next item (its theta was issued in front of this item's second products): lands under the second half's transcendentals
          tc_wait(acc_full(sm, slot ^ 1), ((b + 1u) >> 1) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          tc_ld16_async(ta0 + (slot ^ 1) * kTcbUnits, rA);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t hd[4], rm[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float2 tt = fma2(make_float2(__uint_as_float(rB[8 * c + 2 * v]), __uint_as_float(rB[8 * c + 2 * v + 1])), bc(inv_s), bc(addk));
            const float t0 = tt.x, t1 = tt.y;
            bt_split2((GPODE_BT_EXP & 4) ? t0 : (is_k ? ex2_approx(t0) : __cosf(t0)), (GPODE_BT_EXP & 4) ? t1 : (is_k ? ex2_approx(t1) : __cosf(t1)), hd[v], rm[v]);
          }
          sts128(trow + (2 + c) * 2048, hd[0], hd[1], hd[2], hd[3]);
          sts128(trow + (18 + c) * 2048, rm[0], rm[1], rm[2], rm[3]);
        }
        if (!(GPODE_BT_EXP & 32)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc_arrive(tau_full(sm, slot));   // the issuer may go: theta(b + 2), Q(b), PG(b)
        if (++j == nbi) {
          j = 0;
          ++k;
        }
      }
    }
    sm.blk = b0 + n;
    sm.kk = kk0 + g.D_out;
    sm.pgc = pg0 + static_cast<long>(g.D_out) * nbm;
    __syncthreads();   // every MMA of this evaluation has executed (the last Q commit covers them) and every epilogue is done: xs / X' may change
    __syncthreads();   // (the drain warps write dx between the two barriers)
  }
};

}  // namespace gpode
