// Pipe-peak microbenchmarks for the roofline denominators of the GP vector-field kernels:
// FP32 FMA (scalar FFMA and packed FFMA2), MUFU (ex2 / cos / sin), broadcast LDS.128.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peaks tools/peaks.cu
// Prints one JSON object; all rates are per second over the whole chip, timed with CUDA events.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ILP = 8;

__global__ void k_ffma(float* out, int iters, float a, float b) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678f) out[0] = s;
}

__global__ void k_ffma2(float* out, int iters, float a, float b) {
  float2 acc[ILP];
  float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = __ffma2_rn(acc[i], aa, bb);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i].x + acc[i].y;
  if (s == 12345.678f) out[0] = s;
}

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int OP>
__global__ void k_sfu(float* out, int iters) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3f + i * 0.1f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) acc[i] = ex2a(acc[i]) - 1.0f;   // 1 MUFU + 1 FADD
      if (OP == 1) acc[i] = __cosf(acc[i]);        // FMUL.RZ + MUFU.COS
      if (OP == 2) acc[i] = __sinf(acc[i]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678f) out[0] = s;
}

// every lane reads the same 16-byte word (the access pattern of parameter broadcast)
__global__ void k_lds_bcast(float* out, int iters) {
  __shared__ float4 tab[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = make_float4(i, i + 1, i + 2, i + 3);
  __syncthreads();
  float4 acc = make_float4(0, 0, 0, 0);
  int idx = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      float4 v = tab[(idx + i) & 1023];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    idx += ILP;
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

// The forward inner loop shape: per s-pair 4 broadcast LDS.128, 7 FFMA2, 2 cos (or 2 ex2).
template <int R, bool USE_COS>
__global__ void k_mix(float* out, int iters) {
  __shared__ float4 tab[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = make_float4(1e-3f * i, 1e-3f, 2e-3f, -1e-3f);
  __syncthreads();
  float2 xd[R][6];
  float2 facc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    facc[r] = make_float2(0.f, 0.f);
#pragma unroll
    for (int d = 0; d < 6; ++d) { float v = threadIdx.x * 1e-3f + d + r; xd[r][d] = make_float2(v, v); }
  }
  int idx = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float4 p0 = tab[(idx + 0) & 1023], p1 = tab[(idx + 1) & 1023], p2 = tab[(idx + 2) & 1023], p3 = tab[(idx + 3) & 1023];
      idx += 4;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float2 th = make_float2(p3.x, p3.y);
        th = __ffma2_rn(xd[r][0], make_float2(p0.x, p0.y), th);
        th = __ffma2_rn(xd[r][1], make_float2(p0.z, p0.w), th);
        th = __ffma2_rn(xd[r][2], make_float2(p1.x, p1.y), th);
        th = __ffma2_rn(xd[r][3], make_float2(p1.z, p1.w), th);
        th = __ffma2_rn(xd[r][4], make_float2(p2.x, p2.y), th);
        th = __ffma2_rn(xd[r][5], make_float2(p2.z, p2.w), th);
        float2 c;
        if (USE_COS) { c.x = __cosf(th.x); c.y = __cosf(th.y); }
        else { c.x = ex2a(th.x); c.y = ex2a(th.y); }
        facc[r] = __ffma2_rn(c, make_float2(p3.z, p3.w), facc[r]);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r) s += facc[r].x + facc[r].y;
  if (s == 12345.678f) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  float* out; CK(cudaMalloc(&out, 1024));
  const int iters = 4096, threads = 512, ctas = sms * 4;
  const double nthreads = (double)threads * ctas;
  printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
  double ms;
  ms = time_ms([&] { k_ffma<<<ctas, threads>>>(out, iters, 1.0001f, 0.5f); }, 5);
  printf(", \"ffma_tflops\": %.2f", 2.0 * nthreads * iters * ILP / ms / 1e9);
  ms = time_ms([&] { k_ffma2<<<ctas, threads>>>(out, iters, 1.0001f, 0.5f); }, 5);
  printf(", \"ffma2_tflops\": %.2f", 4.0 * nthreads * iters * ILP / ms / 1e9);
  ms = time_ms([&] { k_sfu<0><<<ctas, threads>>>(out, iters); }, 5);
  printf(", \"ex2_tops\": %.3f", nthreads * iters * ILP / ms / 1e9);
  ms = time_ms([&] { k_sfu<1><<<ctas, threads>>>(out, iters); }, 5);
  printf(", \"cos_tops\": %.3f", nthreads * iters * ILP / ms / 1e9);
  ms = time_ms([&] { k_sfu<2><<<ctas, threads>>>(out, iters); }, 5);
  printf(", \"sin_tops\": %.3f", nthreads * iters * ILP / ms / 1e9);
  ms = time_ms([&] { k_lds_bcast<<<ctas, threads>>>(out, iters); }, 5);
  printf(", \"lds128_bcast_Tinstr_per_s\": %.3f", nthreads / 32 * iters * ILP / ms / 1e9);
  // mix: per iteration 2 s-pairs; per s-pair per state: 14 FMA lanes(7 FFMA2) and 2 SFU
  ms = time_ms([&] { k_mix<1, true><<<ctas, threads>>>(out, iters); }, 5);
  printf(", \"mix_r1_cos_tflops\": %.2f, \"mix_r1_cos_sfu_tops\": %.3f", 2.0 * 14 * 2 * nthreads * iters / ms / 1e9, 2.0 * 2 * nthreads * iters / ms / 1e9);
  ms = time_ms([&] { k_mix<2, true><<<ctas, threads>>>(out, iters); }, 5);
  printf(", \"mix_r2_cos_tflops\": %.2f, \"mix_r2_cos_sfu_tops\": %.3f", 2.0 * 14 * 2 * 2 * nthreads * iters / ms / 1e9, 2.0 * 2 * 2 * nthreads * iters / ms / 1e9);
  ms = time_ms([&] { k_mix<1, false><<<ctas, threads>>>(out, iters); }, 5);
  printf(", \"mix_r1_ex2_tflops\": %.2f, \"mix_r1_ex2_sfu_tops\": %.3f", 2.0 * 14 * 2 * nthreads * iters / ms / 1e9, 2.0 * 2 * nthreads * iters / ms / 1e9);
  ms = time_ms([&] { k_mix<2, false><<<ctas, threads>>>(out, iters); }, 5);
  printf(", \"mix_r2_ex2_tflops\": %.2f, \"mix_r2_ex2_sfu_tops\": %.3f", 2.0 * 14 * 2 * 2 * nthreads * iters / ms / 1e9, 2.0 * 2 * 2 * nthreads * iters / ms / 1e9);
  printf("}\n");
  return 0;
}
