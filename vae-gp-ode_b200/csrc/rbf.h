// Host-side descriptors of the RBF kernels (shared by the kernel translation units and the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "chunk_geom.h"
#include "gpode.h"
#include "sweep_args.h"

namespace gpode {

struct RbfGeom {
  int L, N, NL;          // samples, states per sample, L*N
  int D_in, D_out, DP;   // DP = D_in rounded up to even
  int M, S, MP2, SP2;    // pairs (rounded up)
  int NCs, NCm, RCs, RCm;  // chunks / rows per chunk of the feature and inducing sections
  int stage_floats;      // floats of one pipeline stage (largest chunk)
  int hdr_floats;        // floats per (l,k) header
  int row_floats;        // floats per row
  int order, off;        // ODE order; off = D_in - D_out: where f sits inside the state derivative
  int flags;             // GPODE_FLAG_* of the problem (kernel selection for tests / A-B runs)
  ChunkGeom cg;          // the same chunking, in the form the pipeline reads
};
// packed = [L][D_out][hdr_floats] headers, [L][D_out][SP2+MP2][row_floats] rows, then [L][D_out] max |row coefficient| (the
// power-of-two scale of the fp16 tensor-path dot products)
inline size_t rbf_rows_end_floats(const RbfGeom& g) {
  return static_cast<size_t>(g.L) * g.D_out * (g.hdr_floats + static_cast<size_t>(g.SP2 + g.MP2) * g.row_floats);
}
// D > 8 only: behind the maxima, the operand tiles of the tcgen05 forward sweep (rbf_fwd_tc.cuh), [L][D_out][blocks] tiles of
// kTcfRows units: B operand as no-swizzle K-major core matrices (16-byte chunks of 4 consecutive k; chunk c of row r at
// c * rows * 4 + (r / 8) * 32 + (r % 8) * 4 floats): chunks 0-3 TF32 heads of the 16 coefficients, 4-7 their remainders,
// 8 = (offset head, offset remainder, 0, 0), 9 = 0; then the kTcfRows weights.  Feature units first, inducing units after.
constexpr int kTcfRows = 256;
constexpr int kTcfChunks = 10;
constexpr int kTcfBFloats = kTcfChunks * kTcfRows * 4;
constexpr int kTcfTileFloats = kTcfBFloats + kTcfRows;
// Fused reverse sweep + parameter gradients on tcgen05 (rbf_bwd_tc.cuh): behind the forward tiles, [L][D_out][items] tiles of kTcbUnits units
// ("items": feature units first, inducing units after, each section padded to whole items), 24 KB each, two 12 KB halves that stream
// through separate rings.  All operand blocks are no-swizzle core matrices (8 rows x 16 bytes): 16-byte chunk c of unit r sits at
// c * 2048 + (r / 8) * 128 + (r % 8) * 16 bytes of its block.
//   theta half  (kTcbThBytes): G_h | G_l | off -- fp16 head and fp16 remainder of the 16 coefficients x s_k (2 chunks of 8 each: K = 16 of a
//                kind::f16 step), then a kind::tf32 block of K = 8: (off_h, off_l, 0, 0 | 0 ..) x s_k.  s_k is the exact power-of-two
//                block scale of output k (rbf_pow2_scale of the largest |coefficient|; 1 for every shape of the reference); feature
//                offsets carry + pi/2 (cos(theta + pi/2) = -sin theta, the derivative of the forward's cosine).
//   second half (kTcbPBytes) : B operand (N = kTcbQN columns x K = 128 units, bf16, K-major: chunk kc = unit / 8 of column n at
//                kc * kTcbQN * 16 + (n / 8) * 128 + (n % 8) * 16) of Q[state][n] = sum_u tau[state][u] B[n][u]:
//                n < 16: head of P[u][d = n] = weight_u x coefficient_ud, 16 <= n < 32: its remainder, n = 32 / 33: head / remainder of the
//                weight (inducing units only: the sum the A_k(x) term needs), rest 0.
constexpr int kTcbUnits = 128;
constexpr int kTcbQN = 48;
constexpr int kTcbThBytes = 3 * 4096;
constexpr int kTcbPBytes = kTcbQN * kTcbUnits * 2;
constexpr int kTcbThFloats = kTcbThBytes / 4;
constexpr int kTcbPFloats = kTcbPBytes / 4;
constexpr int kTcbTileFloats = kTcbThFloats + kTcbPFloats;
// exact power-of-two scale s (and 1 / s) of a block of fp16 operands with largest magnitude m: brings m into [2^13, 2^14), the top of
// the fp16 range (products accumulate in fp32), so that the fp16 REMAINDER of every element down to 2^-15 of the maximum is still a
// normal number -- head + remainder then carry 22 bits.  m = 0 (or subnormal): 1.
__host__ __device__ inline void rbf_pow2_scale(float m, float& s, float& inv) {
  union { float f; unsigned u; } c;
  c.f = m;
  const unsigned be = (c.u >> 23) & 255u;
  unsigned bs = be == 0u ? 127u : 267u - be;
  bs = bs > 253u ? 253u : bs;
  c.u = bs << 23;
  s = c.f;
  c.u = (254u - bs) << 23;
  inv = c.f;
}
__host__ __device__ inline int rbf_tc_blocks_s(const RbfGeom& g) { return (g.S + kTcfRows - 1) / kTcfRows; }
__host__ __device__ inline int rbf_tc_blocks(const RbfGeom& g) { return rbf_tc_blocks_s(g) + (g.M + kTcfRows - 1) / kTcfRows; }
inline size_t rbf_tc_floats(const RbfGeom& g) { return g.DP > 8 ? static_cast<size_t>(g.L) * g.D_out * rbf_tc_blocks(g) * kTcfTileFloats : 0; }
__host__ __device__ inline int rbf_tcb_items_s(const RbfGeom& g) { return (g.S + kTcbUnits - 1) / kTcbUnits; }
__host__ __device__ inline int rbf_tcb_items(const RbfGeom& g) { return rbf_tcb_items_s(g) + (g.M + kTcbUnits - 1) / kTcbUnits; }
inline size_t rbf_tcb_floats(const RbfGeom& g) { return g.DP > 8 ? static_cast<size_t>(g.L) * g.D_out * rbf_tcb_items(g) * kTcbTileFloats : 0; }
inline size_t rbf_maxabs_floats(const RbfGeom& g) { return (static_cast<size_t>(g.L) * g.D_out + 3) / 4 * 4; }
inline size_t rbf_packed_floats(const RbfGeom& g) { return rbf_rows_end_floats(g) + rbf_maxabs_floats(g) + rbf_tc_floats(g) + rbf_tcb_floats(g); }
__host__ __device__ inline const float* rbf_tc_tiles_ptr(const float* packed, const RbfGeom& g, int l) {
  return packed + static_cast<size_t>(g.L) * g.D_out * (g.hdr_floats + static_cast<size_t>(g.SP2 + g.MP2) * g.row_floats) +
         (static_cast<size_t>(g.L) * g.D_out + 3) / 4 * 4 + static_cast<size_t>(l) * g.D_out * rbf_tc_blocks(g) * kTcfTileFloats;
}
__host__ __device__ inline const float* rbf_tcb_tiles_ptr(const float* packed, const RbfGeom& g, int l) {
  return rbf_tc_tiles_ptr(packed, g, 0) + static_cast<size_t>(g.L) * g.D_out * rbf_tc_blocks(g) * kTcfTileFloats +
         static_cast<size_t>(l) * g.D_out * rbf_tcb_items(g) * kTcbTileFloats;
}
__host__ __device__ inline const float* rbf_maxabs_ptr(const float* packed, const RbfGeom& g, int l) {
  return packed + static_cast<size_t>(g.L) * g.D_out * (g.hdr_floats + static_cast<size_t>(g.SP2 + g.MP2) * g.row_floats) + static_cast<size_t>(l) * g.D_out;
}
__host__ __device__ inline const float* rbf_hdr_ptr(const float* packed, const RbfGeom& g, int l) {
  return packed + static_cast<size_t>(l) * g.D_out * g.hdr_floats;
}
__host__ __device__ inline const float* rbf_rows_ptr(const float* packed, const RbfGeom& g, int l) {
  return packed + static_cast<size_t>(g.L) * g.D_out * g.hdr_floats + static_cast<size_t>(l) * g.D_out * (g.SP2 + g.MP2) * g.row_floats;
}

struct RbfAccum {        // fp32 accumulators in the workspace, zeroed before the backward
  float* dnu;            // [L][D_out][2*MP2]          sum_n g E
  float* pg;             // [L][D_out][2*MP2][DP]      sum_n g E x_d
  float* dell_x;         // [D_out][DP]                sum_n x_d * dx_k,d
  float* dvar;           // [D_out]                    sum_n g (f - f_p/2)
  float* dell_z;         // [D_out][DP]                sum_m z_d * dZ_k,m,d   (filled by the finalize kernel)
};

using RbfFieldFwdArgs = FieldFwdArgsT<RbfGeom>;
using RbfRolloutFwdArgs = RolloutFwdArgsT<RbfGeom>;
using RbfRolloutBwdArgs = RolloutBwdArgsT<RbfGeom, RbfAccum>;
using RbfFieldBwdArgs = FieldBwdArgsT<RbfGeom, RbfAccum>;

struct RbfPgradArgs {
  RbfGeom g;
  const float* packed;
  const float* xsave;
  const float* gsave;
  long n_te;             // number of (step, stage) slabs
  int chunks;            // CTAs along the state-evaluation axis
  RbfAccum acc;
};

struct RbfPackArgs {
  RbfGeom g;
  int variant;           // GPODE_RBF_SHARED or GPODE_RBF_DIMWISE
  const float* Z;
  const float* ell;
  const float* var;
  const float* eps;
  const float* phase;
  const float* w;
  const float* nu;
  float* packed;
  int with_tc;           // bit 0: also lay out the operand tiles of the tensor-memory forward (forward entry points); bit 1: those of the
                         // tensor-memory reverse sweep / parameter gradients (backward entry points)
};

struct RbfFinalizeArgs {
  RbfGeom g;
  int variant;
  const float* Z;
  const float* ell;
  const float* var;
  const float* nu;
  RbfAccum acc;
  float* d_Z;
  float* d_ell;
  float* d_var;
  float* d_nu;
};

constexpr int kPgThreads = 128;
// block size / pairs per thread / m-blocks of the parameter-gradient kernel
inline void rbf_pgrad_shape(const RbfGeom& g, int& threads, int& PP, int& n_mblk) {
  PP = 1;   // 2 pairs per thread only paid at D = 16 (+4 %), where the tensor-path kernel is used instead
  const int want = (g.MP2 + PP - 1) / PP;                       // threads needed to cover all pairs
  threads = want >= kPgThreads ? kPgThreads : (want + 31) / 32 * 32;
  n_mblk = (g.MP2 + threads * PP - 1) / (threads * PP);
}

// The parameter gradients run on the tensor path (3xTF32 mma.sync, rbf_pgrad_mma.cuh) for D > 8 and on the FFMA path
// (k_rbf_pgrad) for D <= 8 -- measured on B200: 40.3 vs 42.2 ms at D = 16, 0.97 vs 0.79 ms at D = 6.
inline bool rbf_pgrad_use_mma(const RbfGeom& g) { return g.DP > 8; }
// the tensor-path sweep kernels (D > 8) need a chip-filling batch; below it the FFMA kernels run (latency bound there)
inline bool rbf_fwd_use_mma(const RbfGeom& g) { return static_cast<long>(g.N) * g.L >= 32768; }
// Forward sweep at D > 8 once the batch fills the chip: tcgen05 / tensor-memory kernel (rbf_fwd_tc.cuh; 17.1 vs 19.0 ms at config-5
// shapes) unless its 256-unit operand tiles would be more than 15 % padding, in which case (or with GPODE_FLAG_FWD_MMA) the mma.sync
// kernel (RbfMmaFwdPolicy) runs.  Both are parity-tested against the oracle; the choice depends on shapes and the caller's flags only.
inline bool rbf_fwd_use_tc(const RbfGeom& g) {
  if (g.DP <= 8 || static_cast<long>(g.N) * g.L < 32768) return false;
  if (g.flags & GPODE_FLAG_FWD_MMA) return false;
  if (g.flags & GPODE_FLAG_FWD_TCGEN05) return true;    // (tests: force the tensor-memory kernel whatever the padding)
  return static_cast<long>(rbf_tc_blocks(g)) * kTcfRows * 100 <= static_cast<long>(g.S + g.M) * 115;
}
// Reverse sweep at D > 8 on a chip-filling batch: the fused tcgen05 kernel (rbf_bwd_tc.cuh: theta, the transcendental, the state
// gradient AND the parameter-gradient statistics from one tile -- no separate parameter-gradient pass) is the default;
// GPODE_FLAG_BWD_MMA selects the warp-level kernels (RbfMmaBwdPolicy + k_rbf_pgrad_mma).  Both are parity-tested against the oracle.
inline bool rbf_bwd_use_tc(const RbfGeom& g) {
  if (g.DP <= 8 || static_cast<long>(g.N) * g.L < 32768) return false;
  return (g.flags & GPODE_FLAG_BWD_MMA) == 0;
}
// inducing points per CTA of the tensor-path parameter-gradient kernel (8 warps x 16 MT rows)
inline void rbf_pgrad_mma_shape(const RbfGeom& g, int& MT, int& n_mblk) {
  MT = 2 * g.MP2 > 128 ? 2 : 1;
  const int per_cta = 128 * MT;
  n_mblk = (2 * g.MP2 + per_cta - 1) / per_cta;
}
// ---- small-batch policy (rbf_small.cuh): host-side geometry, shared with the ABI's launch-plan query ----
constexpr int kSmWarps = 16;
constexpr int kSmThreads = kSmWarps * 32;
constexpr int kSmStates = 32;
constexpr int kSmMaxCluster = 8;   // portable cluster size
inline int rbf_small_smem_bytes(const RbfGeom& g) {
  return (g.D_out * (g.SP2 + g.MP2) * g.row_floats + g.D_out * g.hdr_floats + 2 * g.DP * kSmStates + 2 * kSmWarps * (g.DP + 2) * kSmStates +
          g.D_out * (g.DP + 1) + g.D_out * kSmStates + 2 * g.D_out * 2 * kSmStates + 2 * kSmMaxCluster * g.DP * kSmStates + 8) * 4;
}
// small batch and a parameter set that fits in shared memory next to the exchange buffers
inline bool rbf_use_small(const RbfGeom& g) {
  return static_cast<long>(g.N) * g.L <= 148L * 32 && rbf_small_smem_bytes(g) <= 200 * 1024;
}
// cluster size of a small-batch launch: as many CTAs per state block as there are outputs (<= 8), while the whole launch fits the chip
inline int rbf_small_cluster(const RbfGeom& g) {
  const long ctas = static_cast<long>((g.N + kSmStates - 1) / kSmStates) * g.L;
  long c = g.D_out < kSmMaxCluster ? g.D_out : kSmMaxCluster;
  if (c * ctas > 148) c = 148 / ctas;
  return c < 1 ? 1 : static_cast<int>(c);
}

// launchers (one per DP instantiation unit); return cudaGetLastError()
cudaError_t rbf_launch_field_fwd(const RbfFieldFwdArgs& a, cudaStream_t st);
cudaError_t rbf_launch_field_bwd(const RbfFieldBwdArgs& a, cudaStream_t st);
cudaError_t rbf_launch_rollout_fwd(const RbfRolloutFwdArgs& a, cudaStream_t st);
cudaError_t rbf_launch_rollout_bwd(const RbfRolloutBwdArgs& a, cudaStream_t st);
cudaError_t rbf_launch_pgrad(const RbfPgradArgs& a, cudaStream_t st);
cudaError_t rbf_launch_pack(const RbfPackArgs& a, cudaStream_t st);
cudaError_t rbf_launch_finalize(const RbfFinalizeArgs& a, cudaStream_t st);
int rbf_smem_bytes(const RbfGeom& g, int threads, int R, bool bwd);  // dynamic shared memory of the sweep kernels

}  // namespace gpode

namespace gpode {
// per-DP instantiations (rbf_inst.cu compiled once per GPODE_DP)
template <int DP> cudaError_t rbf_field_fwd_dp(const RbfFieldFwdArgs& a, cudaStream_t st);
template <int DP> cudaError_t rbf_field_bwd_dp(const RbfFieldBwdArgs& a, cudaStream_t st);
template <int DP> cudaError_t rbf_rollout_fwd_dp(const RbfRolloutFwdArgs& a, cudaStream_t st);
template <int DP> cudaError_t rbf_rollout_bwd_dp(const RbfRolloutBwdArgs& a, cudaStream_t st);
template <int DP> cudaError_t rbf_pgrad_dp(const RbfPgradArgs& a, cudaStream_t st);
}  // namespace gpode
