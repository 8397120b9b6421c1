// Host-side entry points of elbo_kernels.cu: device-side function-sample draws (Philox4x32-10) and the fused, reduced Bernoulli
// log-likelihood of the reconstructions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpode {

cudaError_t philox_fill(int nseg, float* const* outs, const unsigned long long* ns, const int* kinds, unsigned long long seed, unsigned long long offset,
                        cudaStream_t st);
cudaError_t philox_raw(const uint32_t* ctr, const uint32_t* key, uint32_t* out, int n, cudaStream_t st);
cudaError_t bernoulli_forward(int L, int N, long P, const float* z, const float* x, float* lhood, double* acc, cudaStream_t st);
cudaError_t bernoulli_backward(int L, int N, long P, const float* z, const float* x, const float* g, float* dz, cudaStream_t st);

}  // namespace gpode
