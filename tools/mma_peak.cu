// Throughput of the legacy warp-level tensor path (mma.sync m16n8k8 TF32, HMMA) on B200 -- is it fast enough to take
// the D-length dot products of the parameter-gradient recomputation (3xTF32)?  Prints JSON.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_mma(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f400000u}, b[2] = {0x3f800000u, 0x3f100000u + threadIdx.x};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// fp16 / bf16 m16n8k16 (HMMA.16816): twice the K of the TF32 shape per instruction
template <int KIND>
__global__ void k_mma16(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a[4] = {0x3c003c00u + threadIdx.x, 0x38003800u, 0x34003400u, 0x3a003a00u}, b[2] = {0x3c003c00u, 0x39003900u + threadIdx.x};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND>
void mma16(float* out) {
  const int iters = 20000, warps = 16;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_mma16<KIND><<<148, warps * 32>>>(out, 100);
  cudaEventRecord(e0);
  k_mma16<KIND><<<148, warps * 32>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = 148.0 * warps * 8.0 * iters;
  printf("{\"shape\": \"m16n8k16 %s\", \"tflops\": %.1f, \"cycles_per_mma_per_smsp\": %.2f}\n", KIND == 0 ? "f16" : "bf16", mmas * 4096 / (ms * 1e-3) / 1e12,
         (ms * 1e-3 * 1.965e9) / (mmas / (148.0 * 4.0)));
}
template <int CH>
__global__ void k_chain(float* out, int iters) {
  float c[CH][4];
  for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f400000u}, b[2] = {0x3f800000u, 0x3f100000u + threadIdx.x};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
  for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// closer to the kernel: 12 MMAs per "tile" with distinct A / B registers and the accumulator pattern of the 3xTF32 kernel,
// plus ELEM independent FMA-pipe instructions per tile
template <int ELEM>
__global__ void k_tile(float* out, int iters) {
  unsigned a[6][4], b[8][2];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 4; ++j) a[i][j] = 0x3f800000u + threadIdx.x + 16 * i + j;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 2; ++j) b[i][j] = 0x3f000000u + threadIdx.x + 8 * i + j;
  float th[2][4] = {}, pg[4][4] = {}, e[8] = {1.f, 2.f, 3.f, 4.f, 5.f, 6.f, 7.f, 8.f};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int q = 0; q < 6; ++q)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(th[mt][0]), "+f"(th[mt][1]), "+f"(th[mt][2]), "+f"(th[mt][3])
                     : "r"(a[q][0]), "r"(a[q][1]), "r"(a[q][2]), "r"(a[q][3]), "r"(b[q][0]), "r"(b[q][1]));
#pragma unroll
    for (int i = 0; i < ELEM; ++i) e[i & 7] = fmaf(e[i & 7], 1.0001f, 0.5f);
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int q = 0; q < 3; ++q)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(pg[c][0]), "+f"(pg[c][1]), "+f"(pg[c][2]), "+f"(pg[c][3])
                     : "r"(a[q + c / 2][0]), "r"(a[q + c / 2][1]), "r"(a[q + 1][2]), "r"(a[q][3]), "r"(b[q + c][0]), "r"(b[q + c][1]));
  }
  float s = 0.f;
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) s += th[i][j];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += pg[i][j];
  for (int i = 0; i < 8; ++i) s += e[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ELEM>
void tile(float* out, int warps) {
  const int iters = 5000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_tile<ELEM><<<148, warps * 32>>>(out, 100);
  cudaEventRecord(e0);
  k_tile<ELEM><<<148, warps * 32>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double per_smsp = ms * 1e-3 * 1.965e9 / (double(iters) * 24) / (warps / 4.0);
  printf("{\"tile_pattern_elem_instr\": %d, \"warps_per_sm\": %d, \"cycles_per_mma_per_smsp\": %.2f}\n", ELEM, warps, per_smsp);
}
template <int CH>
void chain(float* out, int warps) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_chain<CH><<<148, warps * 32>>>(out, 100);
  cudaEventRecord(e0);
  k_chain<CH><<<148, warps * 32>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double per_warp = ms * 1e-3 * 1.965e9 / (double(iters) * CH);
  printf("{\"dependent_chains_per_warp\": %d, \"warps_per_sm\": %d, \"cycles_per_mma_per_warp\": %.2f, \"cycles_per_mma_per_smsp\": %.2f}\n", CH, warps, per_warp,
         per_warp / (warps / 4.0));
}
int main() {
  {
    float* o; cudaMalloc(&o, 148 * 1024 * 4);
    mma16<0>(o); mma16<1>(o);
    tile<0>(o, 16); tile<64>(o, 16); tile<128>(o, 16); tile<128>(o, 8); tile<256>(o, 16);
    chain<1>(o, 4); chain<2>(o, 4); chain<4>(o, 4); chain<8>(o, 4); chain<1>(o, 16); chain<2>(o, 16); chain<4>(o, 16);
  }
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mma<<<148 * 2, warps * 32 / 2>>>(out, 100);
    cudaEventRecord(e0);
    k_mma<<<148 * 2, warps * 32 / 2>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = 148.0 * warps * 8.0 * iters;   // warp-level instructions
    const double per_smsp_clk = (ms * 1e-3 * 1.965e9) / (mmas / (148.0 * 4.0));
    printf("{\"warps_per_sm\": %d, \"mma_m16n8k8_tf32_per_s\": %.4g, \"tflops\": %.1f, \"cycles_per_mma_per_smsp\": %.2f}\n", warps, mmas / (ms * 1e-3),
           mmas * 2048 * 2 / (ms * 1e-3) / 1e12 / 2, per_smsp_clk);
  }
  return 0;
}
