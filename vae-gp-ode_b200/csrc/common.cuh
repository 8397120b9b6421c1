// Shared device helpers for libgpode (sm_100a only).
//  * packed-FP32 math: FFMA2 / FADD2 / FMUL2 (two fp32 lanes per issue slot, Blackwell-only SASS)
//  * MUFU wrappers (ex2 / cos / sin approximations)
//  * parameter-tile pipeline: 1-D TMA bulk copy (cp.async.bulk -> UBLKCP) + mbarrier, double buffered
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "gpode.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgpode is written for sm_100a (B200) only"
#endif

namespace gpode {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kMaxD = 16;          // largest GP input dimension compiled
constexpr int kMaxStages = 4;
constexpr int kSmemLimit = 227 * 1024;

// ---- packed fp32 ------------------------------------------------------------------------------
__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 lo(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi(float4 v) { return make_float2(v.z, v.w); }

// ---- MUFU ---------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 ex2_2(float2 v) { return make_float2(ex2_approx(v.x), ex2_approx(v.y)); }
__device__ __forceinline__ float2 cos_2(float2 v) { return make_float2(__cosf(v.x), __cosf(v.y)); }
__device__ __forceinline__ float2 sin_2(float2 v) { return make_float2(__sinf(v.x), __sinf(v.y)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- mbarrier + bulk async copy (TMA 1-D) -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion counted in bytes on `bar`; bytes % 16 == 0, both 16-B aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Streams the per-(sample, output-dim) parameter tiles of one sample through two shared-memory
// buffers: tile g of the sequence is tile (g mod n_tiles) of the sample.  One elected thread issues
// the copies; everybody waits on the tile's mbarrier.  release() is a CTA-wide barrier, so every
// thread of the CTA must walk the same tile sequence.
struct TilePipe {
  float* buf0;
  uint64_t* bar;
  const float* src;
  uint32_t tile_bytes;
  int tile_floats;
  int n_tiles;
  long total;
  long cur;
  int next_k;  // tile index (mod n_tiles) of the next copy to issue (thread 0 only)

  __device__ __forceinline__ void issue(long g) {
    if (g < total) {
      const int b = static_cast<int>(g & 1);
      mbar_expect_tx(bar + b, tile_bytes);
      bulk_g2s(buf0 + b * tile_floats, src + static_cast<size_t>(next_k) * tile_floats, tile_bytes, bar + b);
      next_k = (next_k + 1 == n_tiles) ? 0 : next_k + 1;
    }
  }
  __device__ __forceinline__ void init(float* smem_tiles, uint64_t* bars, const float* sample_base, int tile_floats_, int n_tiles_,
                                       long total_) {
    buf0 = smem_tiles;
    bar = bars;
    src = sample_base;
    tile_floats = tile_floats_;
    tile_bytes = static_cast<uint32_t>(tile_floats_) * 4u;
    n_tiles = n_tiles_;
    total = total_;
    cur = 0;
    next_k = 0;
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      mbar_init(bar + 1, 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      issue(0);
      issue(1);
    }
  }
  __device__ __forceinline__ const float* acquire() {
    const int b = static_cast<int>(cur & 1);
    mbar_wait(bar + b, static_cast<uint32_t>((cur >> 1) & 1));
    return buf0 + b * tile_floats;
  }
  __device__ __forceinline__ void release() {
    __syncthreads();  // every warp is done reading buf[cur & 1]
    if (threadIdx.x == 0) issue(cur + 2);
    ++cur;
  }
};

// ---- packed RBF tile geometry (floats) ----------------------------------------------------------
// tile(l,k) = [ header: c_kd (DP floats, padded to a multiple of 4) |
//               S/2 rows: {omega(2j,d),omega(2j+1,d)}_d<DP, {b,b}, {w',w'} |
//               M/2 rows: {G(2j,d),G(2j+1,d)}_d<DP, {H,H}, {nu',nu'} ]
// every row is (DP+2) float2 = (DP+2)/2 float4; DP = D_in rounded up to even.
__host__ __device__ constexpr int rbf_hdr_floats(int DP) { return ((DP + 3) / 4) * 4; }
__host__ __device__ constexpr int rbf_row_floats(int DP) { return (DP + 2) * 2; }
__host__ __device__ inline int rbf_tile_floats(int DP, int SP2, int MP2) { return rbf_hdr_floats(DP) + (SP2 + MP2) * rbf_row_floats(DP); }

// Butcher tableaux of the fixed-grid methods (strictly lower A, weights b), see oracle/solvers.py
struct Tableau {
  int stages;
  float a[kMaxStages][kMaxStages];
  float b[kMaxStages];
};
__host__ __device__ inline Tableau make_tableau(int method) {
  Tableau t = {};
  if (method == GPODE_EULER) {
    t.stages = 1;
    t.b[0] = 1.f;
  } else if (method == GPODE_MIDPOINT) {
    t.stages = 2;
    t.a[1][0] = 0.5f;
    t.b[1] = 1.f;
  } else {
    t.stages = 4;
    t.a[1][0] = 1.f / 3.f;
    t.a[2][0] = -1.f / 3.f;
    t.a[2][1] = 1.f;
    t.a[3][0] = 1.f;
    t.a[3][1] = -1.f;
    t.a[3][2] = 1.f;
    t.b[0] = 0.125f;
    t.b[1] = 0.375f;
    t.b[2] = 0.375f;
    t.b[3] = 0.125f;
  }
  return t;
}
inline int method_stages(int method) { return method == GPODE_EULER ? 1 : (method == GPODE_MIDPOINT ? 2 : 4); }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace gpode
