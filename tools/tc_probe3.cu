// Probe of the tcgen05 building blocks of the FUSED reverse sweep + parameter gradients (sm_100a): the weighted transcendental tile tau
// (128 states x 128 units, bf16 head plane + bf16 remainder plane) lives in SHARED memory in ONE layout that is read two ways:
//   Q  [128 states x N] = tau   B1^T   A = tau K-major   (rows = states, k = plane * 128 + unit), 16 k-steps of kind::f16 (K = 16)
//   PG [128 units  x N] = tau^T B2^T   A = tau MN-major  (rows = units,  k = states), per plane 8 k-steps
// tile layout: 16-byte chunk (state s, plane p, unit group c8 = units 8 c8 .. 8 c8 + 7) at (p * 16 + c8) * 2048 + (s / 8) * 128 + (s % 8) * 16.
//   1. numerics of both products (B K-major; for PG also B MN-major = [state][n]) against a double-precision host reference;
//   2. cost per instruction group with clock64 around back-to-back issues from ONE thread (every CTA of a 148-CTA grid), optionally
//      while 16 other warps stream st.shared.v4 into another region (the epilogue's tau stores) -- reports cycles and the writers' bytes.
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
// no-swizzle descriptor: LBO = byte step along K between core matrices (K-major: next 16-byte k chunk; MN-major: next 8 k), SBO = byte step
// along M / N between core matrices
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
__host__ __device__ inline uint32_t idesc_tf32(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24); }
__host__ __device__ inline uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p, e;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p, e;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory"); }

constexpr int NQ = 48;
constexpr int TILE_BYTES = 65536;            // tau tile
constexpr int B1_BYTES = NQ * 256 * 2;       // Q operand, K-major: chunk (n, kc) at kc * 768 + (n / 8) * 128 + (n % 8) * 16
constexpr int B2_BYTES = NQ * 128 * 2;       // PG operand (k = states): K-major chunk (n, kc) at kc * 768 + ...; MN-major chunk (s, n8) at n8 * 2048 + (s / 8) * 128 + (s % 8) * 16
constexpr int TH_BYTES = 2 * 128 * 56 * 4;   // theta operands (timing only)
constexpr int NV = 10;

__device__ __forceinline__ void read_acc(uint32_t tm, int warp, int tid, int ncols, float* out) {
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  for (int c0 = 0; c0 < ncols; c0 += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tm + lane_base + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[tid * ncols + c0 + j] = __uint_as_float(r[j]);
  }
}

__global__ void __launch_bounds__(544) k_probe3(const uint4* tile_g, const uint4* b1_g, const uint4* b2k_g, const uint4* b2m_g, float* Dout, long long* timing, int* status, int reps,
                                                int variant, int writers) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* sTile = sm;
  unsigned char* sB1 = sTile + TILE_BYTES;
  unsigned char* sB2k = sB1 + B1_BYTES;
  unsigned char* sB2m = sB2k + B2_BYTES;
  unsigned char* sTh = sB2m + B2_BYTES;
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int done_flag;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: the role branches below may use the uniform datapath
  for (int i = tid; i < TILE_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(sTile)[i] = tile_g[i];
  for (int i = tid; i < B1_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(sB1)[i] = b1_g[i];
  for (int i = tid; i < B2_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(sB2k)[i] = b2k_g[i];
  for (int i = tid; i < B2_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(sB2m)[i] = b2m_g[i];
  for (int i = tid; i < TH_BYTES / 4; i += blockDim.x) reinterpret_cast<float*>(sTh)[i] = 0.001f * static_cast<float>(i % 97);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    done_flag = 0;
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
  const uint32_t dQ = tmem, dPGk = tmem + 64, dPGm = tmem + 128, dTh = tmem + 256, dPG64 = tmem + 192;
  const uint32_t tile = smem_u32(sTile), b1 = smem_u32(sB1), b2k = smem_u32(sB2k), b2m = smem_u32(sB2m), th = smem_u32(sTh);
  // Q: k-step j covers k chunks 2 j, 2 j + 1 of the tile (plane j / 8, units 16 (j % 8) ..)
  auto issue_q = [&](uint32_t d, int n_hi, int n_lo) {
    const uint32_t id_hi = idesc_bf16(128, n_hi, 0, 0), id_lo = idesc_bf16(128, n_lo, 0, 0);
#pragma unroll
    for (int j = 0; j < 16; ++j) mma_ss_bf16(d, make_desc(tile + j * 2 * 2048, 2048, 128), make_desc(b1 + j * 2 * 768, 768, 128), j < 8 ? id_hi : id_lo, j > 0);
  };
  // PG: per plane p, k-step i covers states 16 i .. 16 i + 15: A = tau^T MN-major (LBO = 128: next 8 states, SBO = 2048: next 8 units)
  auto issue_pg = [&](uint32_t d, int b_mn, int M, int mhalf) {
    const uint32_t id = idesc_bf16(M, NQ, 1, b_mn);
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint64_t da = make_desc(tile + p * 16 * 2048 + mhalf * 8 * 2048 + i * 2 * 128, 128, 2048);
        const uint64_t db = b_mn ? make_desc(b2m + i * 2 * 128, 128, 2048) : make_desc(b2k + i * 2 * 768, 768, 128);
        mma_ss_bf16(d, da, db, id, (p > 0 || i > 0) ? 1u : 0u);
      }
  };
  auto issue_theta_tf32 = [&]() {
    const uint32_t id = idesc_tf32(128, 128);
#pragma unroll
    for (int s = 0; s < 7; ++s) mma_ss_tf32(dTh, make_desc(th + s * 2 * 128 * 16, 128 * 16, 128), make_desc(th + 128 * 56 * 4 + s * 2 * 128 * 16, 128 * 16, 128), id, s > 0);
  };
  auto issue_theta_f16 = [&]() {   // 4 k-steps of K = 16 (operand values irrelevant: timing only)
    const uint32_t id = idesc_bf16(128, 128, 0, 0);
#pragma unroll
    for (int s = 0; s < 4; ++s) mma_ss_bf16(dTh, make_desc(th + s * 2 * 2048, 2048, 128), make_desc(th + 128 * 56 * 4 + s * 2 * 2048, 2048, 128), id, s > 0);
  };
  uint32_t par0 = 0;
  if (variant < 0) {
    // ---- numerics ----
    if (warp == 0) {
      issue_q(dQ, NQ, NQ);
      issue_pg(dPGk, 0, 128, 0);
      issue_pg(dPGm, 1, 128, 0);
      issue_pg(dPG64, 1, 64, 0);          // M = 64: units 0..63 -> lanes 0..63?  (read back all 128 lanes to see where they land)
      commit(&bar[0]);
    }
    const bool ok = mbar_wait_bounded(&bar[0], par0);
    par0 ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok && tid == 0) status[0] = 1;
    if (ok && tid < 128 && blockIdx.x == 0) {
      read_acc(dQ, warp, tid, NQ, Dout);
      read_acc(dPGk, warp, tid, NQ, Dout + 128 * NQ);
      read_acc(dPGm, warp, tid, NQ, Dout + 2 * 128 * NQ);
      read_acc(dPG64, warp, tid, NQ, Dout + 3 * 128 * NQ);
    }
  } else if (warp == 0) {
    // ---- timing (the whole warp runs the issue loop convergently; one elected lane issues) ----
    uint32_t par1 = 0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      switch (variant) {
        case 0: issue_theta_tf32(); break;
        case 1: issue_q(dQ, NQ, NQ); break;
        case 2: issue_q(dQ, NQ, 16); break;
        case 3: issue_pg(dPGm, 1, 128, 0); break;
        case 4: issue_pg(dPGk, 0, 128, 0); break;
        case 5: issue_pg(dPGm, 1, 64, 0); issue_pg(dPG64, 1, 64, 1); break;
        case 6: issue_theta_f16(); break;
        case 7: issue_theta_tf32(); issue_q(dQ, NQ, 16); issue_pg(dPGm, 1, 128, 0); break;
        case 8: issue_theta_f16(); issue_q(dQ, NQ, 16); issue_pg(dPGm, 1, 128, 0); break;
        case 9: issue_q(dQ, 32, 16); break;
      }
    }
    const long long t1 = clock64();
    commit(&bar[1]);
    const bool done = mbar_wait_bounded(&bar[1], par1);
    const long long t2 = clock64();
    done_flag = 1;
    if (!done && tid == 0) status[0] = 2 + variant;
    if (blockIdx.x == 0 && tid == 0) {
      timing[0] = (t2 - t0) / reps;
      timing[1] = (t1 - t0) / reps;
    }
  } else if (writers && tid >= 32) {
    // the epilogue's share of the shared-memory / tensor-memory ports: writers & 1: every warp streams conflict-free 512-byte st.shared.v4 rows
    // into the theta region; writers & 2: every warp streams tcgen05.ld x16 of its lane quarter (64 B per thread per load)
    long long n = 0;
    const uint32_t base = th + ((tid - 32) & 511) * 16;
    const uint32_t taddr = tmem + 256 + ((((tid - 32) >> 5) >> 2) * 32) + (static_cast<uint32_t>(((tid >> 5) & 3) * 32) << 16);
    uint32_t sink = 0;
    while (!done_flag) {
      if (writers & 1) {
#pragma unroll
        for (int j = 0; j < 6; ++j) asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + j * 8192), "r"(static_cast<uint32_t>(n)) : "memory");
      }
      if (writers & 2) {
        uint32_t r[16], q[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                       "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]),
                       "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
                     : "r"(taddr + 16));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) sink ^= r[j] ^ q[j];
      }
      ++n;
    }
    if (sink == 0x12345u) status[1] = 1;
    if (blockIdx.x == 0 && tid == 32) timing[2] = n * ((writers & 1) ? 6 * 16 : 0) * 512 + n * ((writers & 2) ? 128 : 0) * 512;   // bytes moved by all 512 threads (same loop)
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }
static uint16_t bf16_bits(float x) { __nv_bfloat16 b = __float2bfloat16(x); uint16_t u; memcpy(&u, &b, 2); return u; }

int main() {
  srand(5);
  auto rnd = [] { return static_cast<float>(rand()) / RAND_MAX * 2.f - 1.f; };
  // tau planes (values exactly representable in bf16), B1 [n][k = p * 128 + u], B2 [n][s]
  std::vector<float> tau(2 * 128 * 128), B1(NQ * 256), B2(NQ * 128);
  for (int s = 0; s < 128; ++s)
    for (int u = 0; u < 128; ++u) {
      const float t = rnd() * expf(3.f * rnd()), h = bf16_round(t);
      tau[(0 * 128 + s) * 128 + u] = h;
      tau[(1 * 128 + s) * 128 + u] = bf16_round(t - h);
    }
  for (auto& v : B1) v = bf16_round(rnd());
  for (auto& v : B2) v = bf16_round(rnd());
  std::vector<uint16_t> tile(TILE_BYTES / 2), b1(B1_BYTES / 2), b2k(B2_BYTES / 2), b2m(B2_BYTES / 2);
  for (int p = 0; p < 2; ++p)
    for (int s = 0; s < 128; ++s)
      for (int u = 0; u < 128; ++u) tile[((p * 16 + u / 8) * 2048 + (s / 8) * 128 + (s % 8) * 16) / 2 + (u % 8)] = bf16_bits(tau[(p * 128 + s) * 128 + u]);
  for (int n = 0; n < NQ; ++n)
    for (int k = 0; k < 256; ++k) b1[((k / 8) * 768 + (n / 8) * 128 + (n % 8) * 16) / 2 + (k % 8)] = bf16_bits(B1[n * 256 + k]);
  for (int n = 0; n < NQ; ++n)
    for (int s = 0; s < 128; ++s) {
      b2k[((s / 8) * 768 + (n / 8) * 128 + (n % 8) * 16) / 2 + (s % 8)] = bf16_bits(B2[n * 128 + s]);
      b2m[((n / 8) * 2048 + (s / 8) * 128 + (s % 8) * 16) / 2 + (n % 8)] = bf16_bits(B2[n * 128 + s]);
    }
  uint4 *dT, *dB1, *dB2k, *dB2m; float* dD; long long* dTm; int* dS;
  cudaMalloc(&dT, TILE_BYTES); cudaMalloc(&dB1, B1_BYTES); cudaMalloc(&dB2k, B2_BYTES); cudaMalloc(&dB2m, B2_BYTES);
  cudaMalloc(&dD, 4 * 128 * NQ * 4); cudaMalloc(&dTm, 4 * 8); cudaMalloc(&dS, 8);
  cudaMemcpy(dT, tile.data(), TILE_BYTES, cudaMemcpyHostToDevice); cudaMemcpy(dB1, b1.data(), B1_BYTES, cudaMemcpyHostToDevice);
  cudaMemcpy(dB2k, b2k.data(), B2_BYTES, cudaMemcpyHostToDevice); cudaMemcpy(dB2m, b2m.data(), B2_BYTES, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 4 * 128 * NQ * 4); cudaMemset(dS, 0, 8);
  const int smem = TILE_BYTES + B1_BYTES + 2 * B2_BYTES + TH_BYTES + 2048;
  cudaFuncSetAttribute(k_probe3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 64;
  k_probe3<<<148, 544, smem>>>(dT, dB1, dB2k, dB2m, dD, dTm, dS, reps, -1, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) fprintf(stderr, "numerics: %s\n", cudaGetErrorString(e));
  std::vector<float> D(4 * 128 * NQ);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  long long T[NV][4][3];
  memset(T, 0, sizeof(T));
  for (int w = 0; w < 4 && e == cudaSuccess; ++w)
    for (int v = 0; v < NV && e == cudaSuccess; ++v) {
      cudaMemset(dTm, 0, 4 * 8);
      k_probe3<<<148, 544, smem>>>(dT, dB1, dB2k, dB2m, dD, dTm, dS, reps, v, w);
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) fprintf(stderr, "variant %d writers %d: %s\n", v, w, cudaGetErrorString(e));
      cudaMemcpy(T[v][w], dTm, 3 * 8, cudaMemcpyDeviceToHost);
    }
  int st = -1;
  cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
  double eq = 0, epk = 0, epm = 0, ep64 = 0, ep64_hi = 0, mq = 0, mp = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < NQ; ++n) {
      double q = 0, pg = 0;
      for (int p = 0; p < 2; ++p)
        for (int j = 0; j < 128; ++j) {
          q += double(tau[(p * 128 + m) * 128 + j]) * B1[n * 256 + p * 128 + j];      // m = state, j = unit
          pg += double(tau[(p * 128 + j) * 128 + m]) * B2[n * 128 + j];                // m = unit, j = state
        }
      eq = fmax(eq, fabs(q - D[m * NQ + n]));
      epk = fmax(epk, fabs(pg - D[(128 + m) * NQ + n]));
      epm = fmax(epm, fabs(pg - D[(256 + m) * NQ + n]));
      if (m < 64) ep64 = fmax(ep64, fabs(pg - D[(384 + m) * NQ + n]));
      else ep64_hi = fmax(ep64_hi, fabs(D[(384 + m) * NQ + n]));
      mq = fmax(mq, fabs(q));
      mp = fmax(mp, fabs(pg));
    }
  printf("{\"cuda\": \"%s\", \"status\": %d, \"err_Q_kmajor\": %.3g, \"err_PG_Amn_Bk\": %.3g, \"err_PG_Amn_Bmn\": %.3g, \"err_PG_M64_lanes0_63\": %.3g, \"max_abs_M64_lanes64_127\": %.3g, "
         "\"ref_max_Q\": %.3g, \"ref_max_PG\": %.3g,\n \"cycles_per_rep\": {",
         cudaGetErrorString(e), st, eq, epk, epm, ep64, ep64_hi, mq, mp);
  const char* names[NV] = {"theta_7xSS_tf32_N128", "Q_16xSS_N48", "Q_8xN48_8xN16", "PG_16x_Amn_Bmn_N48", "PG_16x_Amn_Bk_N48", "PG_32x_M64", "theta_4xSS_f16_N128",
                           "item_tf32theta_Q_PG", "item_f16theta_Q_PG", "Q_8xN32_8xN16"};
  for (int v = 0; v < NV; ++v)
    printf("\"%s\": {\"alone\": %lld, \"with_sts\": [%lld, %.1f], \"with_ldtm\": [%lld, %.1f], \"with_both\": [%lld, %.1f]}%s", names[v], T[v][0][0], T[v][1][0],
           T[v][1][0] > 0 ? double(T[v][1][2]) / (double(T[v][1][0]) * reps) : 0.0, T[v][2][0], T[v][2][0] > 0 ? double(T[v][2][2]) / (double(T[v][2][0]) * reps) : 0.0, T[v][3][0],
           T[v][3][0] > 0 ? double(T[v][3][2]) / (double(T[v][3][0]) * reps) : 0.0, v + 1 < NV ? ",\n  " : "");
  printf("},\n \"note\": \"cycles per repetition incl. completion [, bytes per cycle moved by the 16 side warps: st.shared.v4 streams / tcgen05.ld x16 streams / both]; 148 CTAs x 544 threads\"}\n");
  return 0;
}
