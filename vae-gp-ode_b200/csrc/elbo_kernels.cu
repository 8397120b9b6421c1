// The two pieces either side of the GP flow that SURVEY.md section 8(f) rank 3 names (sm_100a):
//
//  1. Function-sample draws on the device.  The reference draws w, eps (standard normal) and the phases (uniform) of every function
//     sample, plus the inducing noise, with numpy on the host (experiments/model/core/kernels.py:13-26,126-137 / 305-316,
//     svpy.py:12-18,94) and copies them over: ~58 k numbers = 1.5 ms of host time per rollout at BASELINE config 2, during which the
//     GPU idles.  k_philox_fill: counter-based Philox4x32-10 (Salmon et al., SC'11; the generator of cuRAND / torch.cuda), one counter
//     per 4 outputs, up to four output segments (w, eps, phase, eps_u) in ONE launch; normals by Box-Muller on pairs of 24-bit uniforms.
//     The stream is a pure function of (seed, offset, segment, index): reproducible, restated in oracle/philox.py.
//
//  2. Bernoulli log-likelihood of the reconstructions, reduced.  The reference materialises X.repeat(L), log(z) X + log(1 - z)(1 - X)
//     over (L, N, T, 1, 28, 28) and then sums over (T, pixels) and averages over L (core/vae.py:136-153, create_model.py:51-53): ~20
//     elementwise passes over the largest tensor of the model.  k_bernoulli_fwd reads z once and X once per sample (L2 resident) and
//     leaves lhood[n]; k_bernoulli_bwd writes dL/dz = (x / z - (1 - x) / (1 - z)) g_n / L.  HBM-bound: 4 B (forward) / 8 B (backward)
//     per element of z.  Arithmetic exactly as the reference writes it (x is NORMALISED, not in [0, 1]: SURVEY Appendix B.7), fp32
//     logf, per-trajectory sums accumulated in double.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace gpode {

namespace {

struct PhiloxSeg {
  float* out;
  unsigned long long n;
  int kind;   // 0 standard normal, 1 uniform [0, 1)
};
struct PhiloxArgs {
  PhiloxSeg seg[4];
  int nseg;
  unsigned long long seed, offset;
};

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// counter = (index / 4 + offset  [64 bit], segment, 0), key = seed [64 bit]
__global__ void __launch_bounds__(256) k_philox_fill(const PhiloxArgs a) {
  const int s = blockIdx.y;
  if (s >= a.nseg) return;
  const PhiloxSeg sg = a.seg[s];
  const unsigned long long quads = (sg.n + 3) / 4;
  for (unsigned long long q = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < quads;
       q += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
    const unsigned long long ctr = q + a.offset;
    uint32_t c[4] = {static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), static_cast<uint32_t>(s), 0u};
    philox4x32_10(c, static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32));
    float v[4];
    if (sg.kind == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = static_cast<float>(c[i] >> 8) * 5.9604644775390625e-08f;               // 2^-24: [0, 1)
    } else {
#pragma unroll
      for (int i = 0; i < 4; i += 2) {
        const float u1 = (static_cast<float>(c[i] >> 8) + 1.0f) * 5.9604644775390625e-08f;                      // (0, 1]
        const float u2 = static_cast<float>(c[i + 1] >> 8) * 5.9604644775390625e-08f;
        const float r = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        v[i] = r * cs;
        v[i + 1] = r * sn;
      }
    }
    const unsigned long long e0 = 4 * q;
    if (e0 + 3 < sg.n && (reinterpret_cast<uintptr_t>(sg.out) & 15) == 0) {
      *reinterpret_cast<float4*>(sg.out + e0) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (e0 + i < sg.n) sg.out[e0 + i] = v[i];
    }
  }
}

// raw generator output for the known-answer test: out[4 i ..] = philox4x32_10(counter_i, key)
__global__ void k_philox_raw(const uint32_t* __restrict__ ctr, const uint32_t* __restrict__ key, uint32_t* __restrict__ out, const int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t c[4] = {ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3]};
  philox4x32_10(c, key[2 * i], key[2 * i + 1]);
#pragma unroll
  for (int j = 0; j < 4; ++j) out[4 * i + j] = c[j];
}

// ---- Bernoulli log-likelihood ----------------------------------------------------------------
// z (L, N, P) reconstructions in (0, 1), x (N, P) targets, P = T * pixels: lhood[n] = 1/L sum_{l, p} log(z) x + log(1 - z)(1 - x).
// grid (chunks of P, N): every CTA reduces one chunk of one trajectory over all L samples; double atomics on lhood64[n].
constexpr int kBeThreads = 256;
__global__ void __launch_bounds__(kBeThreads) k_bernoulli_fwd(const int L, const int N, const long P, const float* __restrict__ z, const float* __restrict__ x,
                                                               double* __restrict__ acc, const long chunk) {
  const int n = blockIdx.y;
  const long p0 = static_cast<long>(blockIdx.x) * chunk, p1 = min(P, p0 + chunk);
  double s = 0.0;
  const float* xr = x + static_cast<size_t>(n) * P;
  const bool vec = (P & 3) == 0 && (p0 & 3) == 0;
  if (vec) {
    for (long p = p0 + 4 * threadIdx.x; p < p1; p += 4 * kBeThreads) {
      const float4 xv = *reinterpret_cast<const float4*>(xr + p);
      float part = 0.f;
      for (int l = 0; l < L; ++l) {
        const float4 zv = __ldcs(reinterpret_cast<const float4*>(z + (static_cast<size_t>(l) * N + n) * P + p));     // streamed once
        part += logf(zv.x) * xv.x + logf(1.f - zv.x) * (1.f - xv.x);
        part += logf(zv.y) * xv.y + logf(1.f - zv.y) * (1.f - xv.y);
        part += logf(zv.z) * xv.z + logf(1.f - zv.z) * (1.f - xv.z);
        part += logf(zv.w) * xv.w + logf(1.f - zv.w) * (1.f - xv.w);
      }
      s += static_cast<double>(part);
    }
  } else {
    for (long p = p0 + threadIdx.x; p < p1; p += kBeThreads) {
      const float xv = xr[p];
      float part = 0.f;
      for (int l = 0; l < L; ++l) {
        const float zv = z[(static_cast<size_t>(l) * N + n) * P + p];
        part += logf(zv) * xv + logf(1.f - zv) * (1.f - xv);
      }
      s += static_cast<double>(part);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double red[kBeThreads / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kBeThreads / 32; ++w) t += red[w];
    atomicAdd(&acc[n], t);
  }
}
__global__ void k_bernoulli_out(const int N, const int L, const double* __restrict__ acc, float* __restrict__ lhood) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) lhood[n] = static_cast<float>(acc[n] / L);
}
// dz[l, n, p] = g[n] / L * (x / z - (1 - x) / (1 - z))
__global__ void __launch_bounds__(kBeThreads) k_bernoulli_bwd(const int L, const int N, const long P, const float* __restrict__ z, const float* __restrict__ x,
                                                               const float* __restrict__ g, float* __restrict__ dz) {
  const size_t total = static_cast<size_t>(L) * N * P;
  const float invL = 1.f / static_cast<float>(L);
  if ((P & 3) == 0) {
    for (size_t e = 4 * (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x); e < total; e += 4 * static_cast<size_t>(gridDim.x) * blockDim.x) {
      const size_t ln = e / P;
      const long p = static_cast<long>(e - ln * P);
      const int n = static_cast<int>(ln % N);
      const float4 zv = __ldcs(reinterpret_cast<const float4*>(z + e));
      const float4 xv = *reinterpret_cast<const float4*>(x + static_cast<size_t>(n) * P + p);
      const float gs = g[n] * invL;
      float4 o;
      o.x = gs * (xv.x / zv.x - (1.f - xv.x) / (1.f - zv.x));
      o.y = gs * (xv.y / zv.y - (1.f - xv.y) / (1.f - zv.y));
      o.z = gs * (xv.z / zv.z - (1.f - xv.z) / (1.f - zv.z));
      o.w = gs * (xv.w / zv.w - (1.f - xv.w) / (1.f - zv.w));
      __stcs(reinterpret_cast<float4*>(dz + e), o);
    }
  } else {
    for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const size_t ln = e / P;
      const long p = static_cast<long>(e - ln * P);
      const int n = static_cast<int>(ln % N);
      const float zv = z[e], xv = x[static_cast<size_t>(n) * P + p];
      dz[e] = g[n] * invL * (xv / zv - (1.f - xv) / (1.f - zv));
    }
  }
}

}  // namespace

cudaError_t philox_fill(int nseg, float* const* outs, const unsigned long long* ns, const int* kinds, unsigned long long seed, unsigned long long offset,
                        cudaStream_t st) {
  PhiloxArgs a;
  unsigned long long mx = 0;
  a.nseg = nseg;
  a.seed = seed;
  a.offset = offset;
  for (int i = 0; i < 4; ++i) {
    a.seg[i].out = i < nseg ? outs[i] : nullptr;
    a.seg[i].n = i < nseg ? ns[i] : 0;
    a.seg[i].kind = i < nseg ? kinds[i] : 0;
    if (a.seg[i].n > mx) mx = a.seg[i].n;
  }
  if (mx == 0) return cudaSuccess;
  unsigned long long blocks = ((mx + 3) / 4 + 255) / 256;
  if (blocks > 148ull * 8) blocks = 148ull * 8;
  k_philox_fill<<<dim3(static_cast<unsigned>(blocks), nseg), 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t philox_raw(const uint32_t* ctr, const uint32_t* key, uint32_t* out, int n, cudaStream_t st) {
  k_philox_raw<<<(n + 127) / 128, 128, 0, st>>>(ctr, key, out, n);
  return cudaGetLastError();
}
cudaError_t bernoulli_forward(int L, int N, long P, const float* z, const float* x, float* lhood, double* acc, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(acc, 0, static_cast<size_t>(N) * sizeof(double), st);
  if (e != cudaSuccess) return e;
  // chunks per trajectory so that the grid fills the chip a few times over (N can be 25): ~ 8 CTAs per SM in total, chunk a multiple of 1024
  long per = (148L * 8 + N - 1) / N;
  long chunk = (P + per - 1) / per;
  chunk = (chunk + 1023) / 1024 * 1024;
  const long nchunk = (P + chunk - 1) / chunk;
  k_bernoulli_fwd<<<dim3(static_cast<unsigned>(nchunk), N), kBeThreads, 0, st>>>(L, N, P, z, x, acc, chunk);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  k_bernoulli_out<<<(N + 255) / 256, 256, 0, st>>>(N, L, acc, lhood);
  return cudaGetLastError();
}
cudaError_t bernoulli_backward(int L, int N, long P, const float* z, const float* x, const float* g, float* dz, cudaStream_t st) {
  const size_t total = static_cast<size_t>(L) * N * P;
  size_t blocks = (total / 4 + kBeThreads - 1) / kBeThreads;
  if (blocks > 148u * 16) blocks = 148u * 16;
  if (blocks < 1) blocks = 1;
  k_bernoulli_bwd<<<static_cast<unsigned>(blocks), kBeThreads, 0, st>>>(L, N, P, z, x, g, dz);
  return cudaGetLastError();
}

}  // namespace gpode
