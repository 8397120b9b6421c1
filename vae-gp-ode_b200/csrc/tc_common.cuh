// tcgen05 / tensor-memory helpers shared by the tensor-memory kernels (rbf_fwd_tc.cuh, rbf_bwd_tc.cuh), sm_100a:
// shared-memory matrix descriptors (no-swizzle K-major core matrices), instruction descriptors, MMA issue (operands from shared
// memory or A from tensor memory), commit -> mbarrier, tensor-memory loads / stores, bounded mbarrier waits.
#pragma once

#include "common.cuh"
#include "rbf.h"

namespace gpode {

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
template <class B>
__device__ __forceinline__ void tc_commit(B bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <class B>
__device__ __forceinline__ void tc_wait(B bar, uint32_t parity) { mbar_wait(bar, parity); }   // time-bounded (common.cuh)
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

template <class B>
__device__ __forceinline__ void tc_arrive(B bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

}  // namespace gpode
