// Probe of the tcgen05 building blocks used by the tensor-memory parameter-gradient kernel (sm_100a):
//   1. D1[128 x 64]  = A1[128 x K1] * B1[64 x K1]^T    kind::tf32, A and B from shared memory (no-swizzle K-major core matrices)
//   2. D2[128 x 16] += A2[128 x 64] * B2[16 x 64]^T     kind::tf32, A from tensor memory (written with tcgen05.st), B from shared memory
// Compares with a CPU reference that rounds operands to TF32 (truncation) -- prints max abs errors as JSON.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 22); ++spin) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
// no-swizzle K-major operand: element (row, k) of a [ROWS x K] fp32/tf32 tile; 16-byte K chunks are ROWS*16 bytes apart
__host__ __device__ inline int km_index(int ROWS, int row, int k) { return (k / 4) * (ROWS * 4) + (row / 8) * 32 + (row % 8) * 4 + (k % 4); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;   // descriptor version (Blackwell)
  return d;                              // layout_type = 0 (no swizzle), base_offset = 0
}
__host__ __device__ inline uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
               "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc),
               "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ uint64_t make_desc_full(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) { return make_desc(saddr, lbo_bytes, sbo_bytes); }
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }

constexpr int K1 = 56, N1 = 64, N2 = 16, K2 = 64;

__global__ void __launch_bounds__(128) k_probe(const float* A1g, const float* B1g, const float* A2g, const float* B2g, float* D1, float* D2, float* D3, int* status) {
  extern __shared__ __align__(1024) float sm[];
  float* sA1 = sm;                        // 128 x 56
  float* sB1 = sA1 + 128 * K1;            // 64 x 56
  float* sB2 = sB1 + N1 * K1;             // 16 x 64
  float* sB3 = sB2 + N2 * K2;             // the same 16 x 64 operand stored [e][c] like a K-major [64 x 16] tile (MN-major for product 2)
  __shared__ __align__(8) uint64_t bar[3];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * K1; i += 128) sA1[km_index(128, i / K1, i % K1)] = A1g[i];
  for (int i = tid; i < N1 * K1; i += 128) sB1[km_index(N1, i / K1, i % K1)] = B1g[i];
  for (int i = tid; i < N2 * K2; i += 128) sB2[km_index(N2, i / K2, i % K2)] = B2g[i];
  for (int i = tid; i < N2 * K2; i += 128) sB3[km_index(K2, i % K2, i / K2)] = B2g[i];   // row = e (k index), col = c (n index)
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&bar[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the async (tensor) proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t d1 = tmem, a2 = tmem + 64, d2 = tmem + 128, d3 = tmem + 160;   // columns
  // ---- 1. SS MMA ----
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N1);
    for (int ks = 0; ks < K1 / 8; ++ks) {
      const uint64_t da = make_desc(smem_u32(sA1) + ks * 2 * (128 * 16), 128 * 16, 128);
      const uint64_t db = make_desc(smem_u32(sB1) + ks * 2 * (N1 * 16), N1 * 16, 128);
      mma_ss(d1, da, db, idesc, ks > 0);
    }
    commit(&bar[0]);
  }
  bool ok = mbar_wait_bounded(&bar[0], 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok) { if (tid == 0) status[0] = 1; }
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  if (ok) {
    for (int c0 = 0; c0 < N1; c0 += 16) {
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                     "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(d1 + lane_base + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) D1[tid * N1 + c0 + j] = __uint_as_float(r[j]);
    }
    // ---- 2. write A2 (row = tid) into tensor memory, TS MMA ----
    for (int c0 = 0; c0 < K2; c0 += 16) {
      uint32_t r[16];
      for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(A2g[tid * K2 + c0 + j]);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a2 + lane_base + c0), "r"(r[0]), "r"(r[1]),
                   "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
                   "r"(r[15])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (ok) {
    if (tid == 0) {
      const uint32_t idesc = make_idesc(128, N2);
      for (int ks = 0; ks < K2 / 8; ++ks) {
        const uint64_t db = make_desc(smem_u32(sB2) + ks * 2 * (N2 * 16), N2 * 16, 128);
        mma_ts(d2, a2 + ks * 8, db, idesc, ks > 0);
      }
      commit(&bar[1]);
    }
    ok = mbar_wait_bounded(&bar[1], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) { if (tid == 0) status[0] = 2; }
    if (ok) {
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                     "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(d2 + lane_base));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) D2[tid * N2 + j] = __uint_as_float(r[j]);
    }
    // ---- 3. TS MMA, B MN-major: element (n = c, k = e) at (c/4) * K2*4 + (e/8)*32 + (e%8)*4 + c%4; try both LBO / SBO assignments ----
    for (int variant = 0; variant < 2 && ok; ++variant) {
      const uint32_t d3v = d3 + 16 * variant;
      if (tid == 0) {
        const uint32_t idesc = make_idesc(128, N2) | (1u << 16);   // b_major = MN
        for (int ks = 0; ks < K2 / 8; ++ks) {
          const uint64_t db = variant == 0 ? make_desc(smem_u32(sB3) + ks * 128, 128, K2 * 16) : make_desc(smem_u32(sB3) + ks * 128, K2 * 16, 128);
          mma_ts(d3v, a2 + ks * 8, db, idesc, ks > 0);
        }
        commit(&bar[2]);
      }
      ok = mbar_wait_bounded(&bar[2], variant);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (!ok) { if (tid == 0) status[0] = 3; }
      if (ok) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                       "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(d3v + lane_base));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D3[variant * 128 * N2 + tid * N2 + j] = __uint_as_float(r[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  std::vector<float> A1(128 * K1), B1(N1 * K1), A2(128 * K2), B2(N2 * K2);
  srand(1);
  auto rnd = [] { return static_cast<float>(rand()) / RAND_MAX * 2.f - 1.f; };
  for (auto& v : A1) v = rnd();
  for (auto& v : B1) v = rnd();
  for (auto& v : A2) v = rnd();
  for (auto& v : B2) v = rnd();
  float *dA1, *dB1, *dA2, *dB2, *dD1, *dD2, *dD3; int* dst;
  cudaMalloc(&dA1, A1.size() * 4); cudaMalloc(&dB1, B1.size() * 4); cudaMalloc(&dA2, A2.size() * 4); cudaMalloc(&dB2, B2.size() * 4);
  cudaMalloc(&dD1, 128 * N1 * 4); cudaMalloc(&dD2, 128 * N2 * 4); cudaMalloc(&dD3, 2 * 128 * N2 * 4); cudaMemset(dD3, 0, 2 * 128 * N2 * 4); cudaMalloc(&dst, 4);
  cudaMemcpy(dA1, A1.data(), A1.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB1, B1.data(), B1.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dA2, A2.data(), A2.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB2, B2.data(), B2.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD1, 0, 128 * N1 * 4); cudaMemset(dD2, 0, 128 * N2 * 4); cudaMemset(dst, 0, 4);
  const int smem = (128 * K1 + N1 * K1 + 2 * N2 * K2) * 4 + 1024;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_probe<<<1, 128, smem>>>(dA1, dB1, dA2, dB2, dD1, dD2, dD3, dst);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> D1(128 * N1), D2(128 * N2), D3(2 * 128 * N2); int st = -1;
  cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost); cudaMemcpy(D3.data(), dD3, D3.size() * 4, cudaMemcpyDeviceToHost);
  double e1 = 0, e2 = 0, e1x = 0, e2x = 0, e3 = 0, e3b = 0;
  for (int m = 0; m < 128; ++m) {
    for (int n = 0; n < N1; ++n) {
      double s = 0, sx = 0;
      for (int k = 0; k < K1; ++k) { s += double(tf32_trunc(A1[m * K1 + k])) * tf32_trunc(B1[n * K1 + k]); sx += double(A1[m * K1 + k]) * B1[n * K1 + k]; }
      e1 = fmax(e1, fabs(s - D1[m * N1 + n])); e1x = fmax(e1x, fabs(sx - D1[m * N1 + n]));
    }
    for (int n = 0; n < N2; ++n) {
      double s = 0, sx = 0;
      for (int k = 0; k < K2; ++k) { s += double(tf32_trunc(A2[m * K2 + k])) * tf32_trunc(B2[n * K2 + k]); sx += double(A2[m * K2 + k]) * B2[n * K2 + k]; }
      e2 = fmax(e2, fabs(s - D2[m * N2 + n])); e2x = fmax(e2x, fabs(sx - D2[m * N2 + n])); e3 = fmax(e3, fabs(s - D3[m * N2 + n])); e3b = fmax(e3b, fabs(s - D3[128 * N2 + m * N2 + n]));
    }
  }
  printf("{\"cuda\": \"%s\", \"status\": %d, \"ss_err_vs_tf32_trunc\": %.3g, \"ss_err_vs_fp32\": %.3g, \"ts_err_vs_tf32_trunc\": %.3g, \"ts_err_vs_fp32\": %.3g, \"ts_mn_major_err_lbo128_sbo1024\": %.3g, \"ts_mn_major_err_lbo1024_sbo128\": %.3g}\n", cudaGetErrorString(e), st, e1,
         e1x, e2, e2x, e3, e3b);
  return 0;
}
