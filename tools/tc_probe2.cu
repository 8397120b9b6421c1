// Probe of the tcgen05 building blocks of the tensor-memory REVERSE sweep / parameter gradients (sm_100a):
//   1. numerics of   D[128 x N2] = A[128 x 256] * B[N2 x 256]^T   kind::f16 (bf16 inputs, fp32 accumulate) with A read from TENSOR MEMORY
//      as packed bf16 pairs (one 32-bit column = two consecutive k) written by tcgen05.st, B from shared memory (no-swizzle K-major
//      core matrices of 8 rows x 8 bf16);
//   2. cost per instruction, measured with clock64 around back-to-back issues from ONE thread (every CTA of a 148-CTA grid):
//      theta-like SS kind::tf32 (M 128, N 128, K 8) and the skinny TS kind::f16 (M 128, N 16 / 32 / 48, K 16).
// Prints one JSON object.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 24); ++spin) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
__host__ __device__ inline uint32_t idesc_tf32(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24); }
__host__ __device__ inline uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24); }
__device__ __forceinline__ void mma_ss_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_bf16(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }

constexpr int KB = 256;       // k extent of the TS product (bf16 elements) = 128 packed columns
constexpr int N2MAX = 48;
constexpr int THN = 128, THK = 56;   // theta-like SS product

// results[0..]: cycles per repetition of each timed variant (CTA 0), issue-only cycles
__global__ void __launch_bounds__(160) k_probe2(const uint32_t* Apk, const __nv_bfloat16* Bg, float* Dout, long long* timing, int* status, int reps, int vmask) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~static_cast<uintptr_t>(1023));
  __nv_bfloat16* sB2 = reinterpret_cast<__nv_bfloat16*>(sm);                       // 48 x 256 bf16 = 24 KB
  float* sA1 = reinterpret_cast<float*>(sm + N2MAX * KB * 2);                      // 128 x 56 fp32 = 28 KB
  float* sB1 = sA1 + 128 * THK;                                                    // up to 256 x 56
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  // B2 element (n, k): 16-byte chunk c = k / 8 of row n at c * N2MAX * 16 + (n / 8) * 128 + (n % 8) * 16 bytes, + (k % 8) * 2
  for (int i = tid; i < N2MAX * KB; i += blockDim.x) {
    const int n = i / KB, k = i % KB;
    sB2[((k / 8) * N2MAX * 16 + (n / 8) * 128 + (n % 8) * 16) / 2 + (k % 8)] = Bg[i];
  }
  for (int i = tid; i < 3 * 128 * THK; i += blockDim.x) sA1[i] = 0.001f * static_cast<float>(i % 97);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t aA = tmem, dQ = tmem + 128, dTh = tmem + 256;   // A: 128 packed columns; Q: 48; theta: 128
  // ---- A (row = tid) into tensor memory ----
  if (tid < 128) {
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    for (int c0 = 0; c0 < KB / 2; c0 += 16) {
      uint32_t r[16];
      for (int j = 0; j < 16; ++j) r[j] = Apk[tid * (KB / 2) + c0 + j];
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(aA + lane_base + c0), "r"(r[0]), "r"(r[1]),
                   "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t par0 = 0;
  // ---- 1. numerics: 16 TS MMAs of K = 16 ----
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16(128, N2MAX);
    for (int ks = 0; ks < KB / 16; ++ks) mma_ts_bf16(dQ, aA + ks * 8, make_desc(smem_u32(sB2) + ks * 2 * N2MAX * 16, N2MAX * 16, 128), idesc, ks > 0);
    commit(&bar[0]);
  }
  bool ok = mbar_wait_bounded(&bar[0], par0);
  par0 ^= 1u;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok && tid == 0) status[0] = 1;
  if (ok && tid < 128 && blockIdx.x == 0) {
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    for (int c0 = 0; c0 < N2MAX; c0 += 16) {
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                     "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(dQ + lane_base + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) Dout[tid * N2MAX + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- 2. timing: variants v = 0: theta only (7 SS tf32 N=128); 1..3: TS bf16 only with N = 16 / 32 / 48 (16 per rep); 4: theta + TS N=48;
  //                 5: SS tf32 N = 256 (7 per rep) ----
  if (ok && tid == 0) {
    uint32_t par1 = 0;
    for (int v = 0; v < 6; ++v) {
      if (!((vmask >> v) & 1)) continue;
      const int n2 = v == 1 ? 16 : (v == 2 ? 32 : 48);
      const uint32_t id_th = idesc_tf32(128, v == 5 ? 256 : THN), id_q = idesc_bf16(128, n2);
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (v == 0 || v == 4 || v == 5) {
#pragma unroll
          for (int s = 0; s < 7; ++s)
            mma_ss_tf32(v == 5 ? tmem : dTh, make_desc(smem_u32(sA1) + s * 2 * 128 * 16, 128 * 16, 128), make_desc(smem_u32(sB1) + s * 2 * (v == 5 ? 256 : 128) * 16, (v == 5 ? 256 : 128) * 16, 128), id_th, s > 0);
        }
        if (v >= 1 && v <= 4) {
#pragma unroll
          for (int ks = 0; ks < KB / 16; ++ks) mma_ts_bf16(dQ, aA + ks * 8, make_desc(smem_u32(sB2) + ks * 2 * N2MAX * 16, N2MAX * 16, 128), id_q, ks > 0);
        }
      }
      const long long t1 = clock64();
      commit(&bar[1]);
      const bool done = mbar_wait_bounded(&bar[1], par1);
      par1 ^= 1u;
      const long long t2 = clock64();
      if (!done) { status[0] = 2 + v; break; }
      if (blockIdx.x == 0) {
        timing[2 * v] = (t2 - t0) / reps;
        timing[2 * v + 1] = (t1 - t0) / reps;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  srand(3);
  auto rnd = [] { return static_cast<float>(rand()) / RAND_MAX * 2.f - 1.f; };
  // A: tau (128 x 128 fp32) split into bf16 head (k = 2 r) and bf16 remainder (k = 2 r + 1), packed low half = even k
  std::vector<float> tau(128 * 128), Ph(N2MAX * 128);
  std::vector<uint32_t> Apk(128 * 128);
  std::vector<__nv_bfloat16> B(N2MAX * KB);
  std::vector<float> Aval(128 * KB), Bval(N2MAX * KB);
  for (auto& v : tau) v = rnd() * expf(4.f * rnd());
  for (int m = 0; m < 128; ++m)
    for (int r = 0; r < 128; ++r) {
      const float t = tau[m * 128 + r], h = bf16_round(t), l = bf16_round(t - h);
      __nv_bfloat16 hb = __float2bfloat16(h), lb = __float2bfloat16(l);
      uint16_t hu, lu;
      memcpy(&hu, &hb, 2);
      memcpy(&lu, &lb, 2);
      Apk[m * 128 + r] = static_cast<uint32_t>(hu) | (static_cast<uint32_t>(lu) << 16);
      Aval[m * KB + 2 * r] = h;
      Aval[m * KB + 2 * r + 1] = l;
    }
  for (int n = 0; n < N2MAX; ++n)
    for (int k = 0; k < KB; ++k) {
      const float v = bf16_round(rnd());
      B[n * KB + k] = __float2bfloat16(v);
      Bval[n * KB + k] = v;
    }
  uint32_t* dA; __nv_bfloat16* dB; float* dD; long long* dT; int* dS;
  cudaMalloc(&dA, Apk.size() * 4); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * N2MAX * 4); cudaMalloc(&dT, 16 * 8); cudaMalloc(&dS, 4);
  cudaMemcpy(dA, Apk.data(), Apk.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 128 * N2MAX * 4); cudaMemset(dT, 0, 16 * 8); cudaMemset(dS, 0, 4);
  const int smem = N2MAX * KB * 2 + 3 * 128 * THK * 4 + 2048;
  cudaFuncSetAttribute(k_probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 64;
  cudaError_t e = cudaSuccess;
  for (int v = -1; v < 6 && e == cudaSuccess; ++v) {   // v = -1: numerics only; then one timed variant per launch
    k_probe2<<<148, 160, smem>>>(dA, dB, dD, dT, dS, reps, v < 0 ? 0 : (1 << v));
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) fprintf(stderr, "variant %d: %s\n", v, cudaGetErrorString(e));
  }
  std::vector<float> D(128 * N2MAX); long long T[16]; int st = -1;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(T, dT, sizeof(T), cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
  double err = 0, err_swapped = 0, ref_max = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N2MAX; ++n) {
      double s = 0, sw = 0;
      for (int k = 0; k < KB; ++k) {
        s += double(Aval[m * KB + k]) * Bval[n * KB + k];
        sw += double(Aval[m * KB + (k ^ 1)]) * Bval[n * KB + k];
      }
      err = fmax(err, fabs(s - D[m * N2MAX + n]));
      err_swapped = fmax(err_swapped, fabs(sw - D[m * N2MAX + n]));
      ref_max = fmax(ref_max, fabs(s));
    }
  printf("{\"cuda\": \"%s\", \"status\": %d, \"ts_bf16_err_low_half_even_k\": %.3g, \"ts_bf16_err_swapped\": %.3g, \"ref_max\": %.3g, "
         "\"cycles_per_rep\": {\"theta_7xSS_tf32_N128\": [%lld, %lld], \"ts16_bf16_N16\": [%lld, %lld], \"ts16_bf16_N32\": [%lld, %lld], \"ts16_bf16_N48\": [%lld, %lld], "
         "\"theta_plus_ts_N48\": [%lld, %lld], \"theta_7xSS_tf32_N256\": [%lld, %lld]}, \"note\": \"[total incl. completion, issue only] per repetition; 148 CTAs\"}\n",
         cudaGetErrorString(e), st, err, err_swapped, ref_max, T[0], T[1], T[2], T[3], T[4], T[5], T[6], T[7], T[8], T[9], T[10], T[11]);
  return 0;
}
