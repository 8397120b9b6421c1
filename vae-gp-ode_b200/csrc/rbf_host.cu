// RBF: parameter packing, gradient finalisation and the DP dispatch of the sweep kernels.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "rbf.h"

namespace gpode {

int rbf_smem_bytes(const RbfGeom& g, int threads, int R, bool bwd) {
  int floats = 32 + kPipeStages * g.stage_floats + g.D_out * g.hdr_floats + (g.D_out + 3) / 4 * 4 + g.DP * R * threads;
  if (bwd) floats += g.DP * R * threads + g.D_out * (g.DP + 1) + 4 * threads;   // + per-warp scratch (128 floats) of the tensor-path reverse sweep
  else floats += 4 * threads;   // per-warp result scratch (128 floats) of the tensor-path forward
  return floats * 4;
}

// ---------------------------------------------------------------------------------------------
// pack: reference-layout tensors -> headers + rows per (sample, output dim) (layout in common.cuh)
//   omega = eps / ell (kernels.py:120-124), w' = sqrt(var/S) w (kernels.py:149), nu' = var nu (kernels.py:106-107,180);
//   inducing rows carry ln2 * nu' so that the reverse sweep needs no extra multiply (the forward divides once per output)
// ---------------------------------------------------------------------------------------------
__global__ void k_rbf_pack(const RbfPackArgs a) {
  const RbfGeom& g = a.g;
  const int k = blockIdx.y, l = blockIdx.z;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_rows = g.SP2 + g.MP2;
  if (row > n_rows) return;
  const bool dimwise = a.variant == GPODE_RBF_DIMWISE;
  const int DP = g.DP, D_in = g.D_in, D_out = g.D_out, S = g.S, M = g.M;
  const float var_k = dimwise ? a.var[k] : a.var[0];
  auto ell = [&](int d) { return dimwise ? a.ell[k * D_in + d] : a.ell[d]; };
  auto ckd = [&](int d) { const float e = ell(d); return -0.5f * kLog2e / (e * e); };
  if (row == n_rows) {  // header
    float* hdr = const_cast<float*>(rbf_hdr_ptr(a.packed, g, l)) + k * g.hdr_floats;
    for (int d = 0; d < g.hdr_floats; ++d) hdr[d] = d < D_in ? ckd(d) : 0.f;
    return;
  }
  float2* out = reinterpret_cast<float2*>(const_cast<float*>(rbf_rows_ptr(a.packed, g, l)) +
                                          (static_cast<size_t>(k) * n_rows + row) * g.row_floats);
  if (row < g.SP2) {
    const float amp = sqrtf(var_k / static_cast<float>(S));
    for (int d = 0; d < DP; ++d) {
      float v[2] = {0.f, 0.f};
      for (int h = 0; h < 2; ++h) {
        const int s = 2 * row + h;
        if (s < S && d < D_in) {
          const size_t idx = dimwise ? ((static_cast<size_t>(l) * D_in + d) * S + s) * D_out + k : (static_cast<size_t>(l) * D_in + d) * S + s;
          v[h] = a.eps[idx] / ell(d);
        }
      }
      out[d] = make_float2(v[0], v[1]);
    }
    float b[2] = {0.f, 0.f}, w[2] = {0.f, 0.f};
    for (int h = 0; h < 2; ++h) {
      const int s = 2 * row + h;
      if (s < S) {
        b[h] = dimwise ? a.phase[(static_cast<size_t>(l) * S + s) * D_out + k] : a.phase[static_cast<size_t>(l) * S + s];
        w[h] = amp * a.w[(static_cast<size_t>(l) * S + s) * D_out + k];
      }
    }
    out[DP] = make_float2(b[0], b[1]);
    out[DP + 1] = make_float2(w[0], w[1]);
    float mx = 0.f;
    for (int d = 0; d < DP; ++d) mx = fmaxf(mx, fmaxf(fabsf(out[d].x), fabsf(out[d].y)));
    atomicMax(reinterpret_cast<unsigned int*>(const_cast<float*>(rbf_maxabs_ptr(a.packed, g, l)) + k), __float_as_uint(mx));
  } else {
    const int j = row - g.SP2;
    float H[2] = {0.f, 0.f}, nu[2] = {0.f, 0.f};
    for (int d = 0; d < DP; ++d) {
      float v[2] = {0.f, 0.f};
      for (int h = 0; h < 2; ++h) {
        const int m = 2 * j + h;
        if (m < M && d < D_in) {
          const float c = ckd(d), z = a.Z[m * D_in + d];
          v[h] = -2.f * c * z;
          H[h] = fmaf(c * z, z, H[h]);
        }
      }
      out[d] = make_float2(v[0], v[1]);
    }
    for (int h = 0; h < 2; ++h) {
      const int m = 2 * j + h;
      if (m < M)
        nu[h] = kLn2 * var_k * (dimwise ? a.nu[(static_cast<size_t>(l) * D_out + k) * M + m] : a.nu[(static_cast<size_t>(l) * M + m) * D_out + k]);
    }
    out[DP] = make_float2(H[0], H[1]);
    out[DP + 1] = make_float2(nu[0], nu[1]);
    float mx = 0.f;
    for (int d = 0; d < DP; ++d) mx = fmaxf(mx, fmaxf(fabsf(out[d].x), fabsf(out[d].y)));
    atomicMax(reinterpret_cast<unsigned int*>(const_cast<float*>(rbf_maxabs_ptr(a.packed, g, l)) + k), __float_as_uint(mx));
  }
}

// operand tiles of the tcgen05 forward sweep from the packed rows (layout: rbf.h); thread <-> unit of the tile
__global__ void k_rbf_pack_tc(const RbfPackArgs a) {
  const RbfGeom& g = a.g;
  const int blk = blockIdx.x, k = blockIdx.y, l = blockIdx.z, r = threadIdx.x;
  const int nbs = rbf_tc_blocks_s(g), nb = rbf_tc_blocks(g);
  const bool is_m = blk >= nbs;
  const int unit = (is_m ? blk - nbs : blk) * kTcfRows + r;
  const bool real = unit < (is_m ? g.M : g.S);
  const float* row = rbf_rows_ptr(a.packed, g, l) + (static_cast<size_t>(k) * (g.SP2 + g.MP2) + (is_m ? g.SP2 : 0) + (unit >> 1)) * g.row_floats + (unit & 1);
  float* tile = const_cast<float*>(rbf_tc_tiles_ptr(a.packed, g, l)) + (static_cast<size_t>(k) * nb + blk) * kTcfTileFloats;
  auto at = [&](int c) { return tile + c * kTcfRows * 4 + (r >> 3) * 32 + (r & 7) * 4; };
  auto head = [](float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); };
  for (int c = 0; c < 4; ++c) {
    float v[4], h[4];
    for (int i = 0; i < 4; ++i) {
      const int d = 4 * c + i;
      v[i] = (real && d < g.DP) ? row[2 * d] : 0.f;
      h[i] = head(v[i]);
    }
    *reinterpret_cast<float4*>(at(c)) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(at(4 + c)) = make_float4(v[0] - h[0], v[1] - h[1], v[2] - h[2], v[3] - h[3]);
  }
  const float off = real ? row[2 * g.DP] : 0.f, oh = head(off);
  *reinterpret_cast<float4*>(at(8)) = make_float4(oh, off - oh, 0.f, 0.f);
  *reinterpret_cast<float4*>(at(9)) = make_float4(0.f, 0.f, 0.f, 0.f);
  tile[kTcfBFloats + r] = real ? row[2 * g.DP + 2] : 0.f;
}

// operand tiles of the fused tcgen05 reverse sweep from the packed rows (layout: rbf.h); thread <-> unit of the item
__global__ void k_rbf_pack_tcb(const RbfPackArgs a) {
  const RbfGeom& g = a.g;
  const int item = blockIdx.x, k = blockIdx.y, l = blockIdx.z, r = threadIdx.x;
  const int nbs = rbf_tcb_items_s(g), nbi = rbf_tcb_items(g);
  const bool is_m = item >= nbs;
  const int unit = (is_m ? item - nbs : item) * kTcbUnits + r;
  const bool real = unit < (is_m ? g.M : g.S);
  const float* row = rbf_rows_ptr(a.packed, g, l) + (static_cast<size_t>(k) * (g.SP2 + g.MP2) + (is_m ? g.SP2 : 0) + (unit >> 1)) * g.row_floats + (unit & 1);
  unsigned char* tile = reinterpret_cast<unsigned char*>(const_cast<float*>(rbf_tcb_tiles_ptr(a.packed, g, l)) + (static_cast<size_t>(k) * nbi + item) * kTcbTileFloats);
  float sk, inv_sk;
  rbf_pow2_scale(rbf_maxabs_ptr(a.packed, g, l)[k], sk, inv_sk);
  const int core = (r >> 3) * 128 + (r & 7) * 16;   // this unit's 16-byte slot inside a chunk
  float coef[16];
  for (int d = 0; d < 16; ++d) coef[d] = (real && d < g.DP) ? row[2 * d] : 0.f;
  // theta half: fp16 heads / remainders of the scaled coefficients (two chunks of 8 dims each), then the tf32 offset block
  for (int c = 0; c < 2; ++c) {
    __half hh[8], ll[8];
    for (int i = 0; i < 8; ++i) {
      const float v = coef[8 * c + i] * sk;
      hh[i] = __float2half_rn(v);
      ll[i] = __float2half_rn(v - __half2float(hh[i]));
    }
    *reinterpret_cast<uint4*>(tile + c * 2048 + core) = *reinterpret_cast<const uint4*>(hh);
    *reinterpret_cast<uint4*>(tile + 4096 + c * 2048 + core) = *reinterpret_cast<const uint4*>(ll);
  }
  // feature units: cos(theta + pi/2) = -sin(theta), the derivative of the forward's cosine; inducing units: the exponent offset H
  const float off = (real ? row[2 * g.DP] + (is_m ? 0.f : 1.57079632679489662f) : 0.f) * sk;
  const float oh = __uint_as_float(__float_as_uint(off) & 0xFFFFE000u);
  *reinterpret_cast<float4*>(tile + 8192 + core) = make_float4(oh, off - oh, 0.f, 0.f);
  *reinterpret_cast<float4*>(tile + 8192 + 2048 + core) = make_float4(0.f, 0.f, 0.f, 0.f);
  // second half: column n of unit r (k index of the MMA = unit)
  const float wgt = real ? row[2 * g.DP + 2] : 0.f;
  __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(tile + kTcbThBytes);
  auto at = [&](int n) -> __nv_bfloat16& { return P[((r >> 3) * kTcbQN * 16 + (n >> 3) * 128 + (n & 7) * 16) / 2 + (r & 7)]; };
  for (int n = 0; n < kTcbQN; ++n) at(n) = __float2bfloat16_rn(0.f);
  for (int d = 0; d <= 16; ++d) {
    const float pv = d < 16 ? wgt * coef[d] : (is_m ? wgt : 0.f);
    const __nv_bfloat16 ph = __float2bfloat16_rn(pv);
    const __nv_bfloat16 pl = __float2bfloat16_rn(pv - __bfloat162float(ph));
    if (d < 16) {
      at(d) = ph;
      at(16 + d) = pl;
    } else {
      at(32) = ph;
      at(33) = pl;
    }
  }
}

cudaError_t rbf_launch_pack(const RbfPackArgs& a, cudaStream_t st) {
  cudaError_t e0 = cudaMemsetAsync(const_cast<float*>(rbf_maxabs_ptr(a.packed, a.g, 0)), 0, static_cast<size_t>(a.g.L) * a.g.D_out * 4, st);
  if (e0 != cudaSuccess) return e0;
  const int rows = a.g.SP2 + a.g.MP2 + 1;
  dim3 grid((rows + 127) / 128, a.g.D_out, a.g.L);
  k_rbf_pack<<<grid, 128, 0, st>>>(a);
  if ((a.with_tc & 1) && rbf_fwd_use_tc(a.g)) {
    dim3 gt(rbf_tc_blocks(a.g), a.g.D_out, a.g.L);
    k_rbf_pack_tc<<<gt, kTcfRows, 0, st>>>(a);
  }
  if ((a.with_tc & 2) && rbf_bwd_use_tc(a.g)) {
    dim3 gt(rbf_tcb_items(a.g), a.g.D_out, a.g.L);
    k_rbf_pack_tcb<<<gt, kTcbUnits, 0, st>>>(a);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// finalize: accumulators -> gradients in the reference layouts
//   dZ_kmd  = nu'_km / ell_kd^2 (pg_kmd - z_md dnu'_km)            (per output dim, per sample)
//   d_ell_kd = -(1/ell_kd) (sum_n x_nd dx_nkd + sum_m z_md dZ_kmd)  (scale invariance of f_k in (x_d, z_d, ell_kd))
//   d_var_k  = (1/var_k) sum_n g_nk (f_nk - f_p,nk / 2) ;  d_nu_km = var_k dnu'_km
// ---------------------------------------------------------------------------------------------
__global__ void k_rbf_finalize_z(const RbfFinalizeArgs a) {
  const RbfGeom& g = a.g;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.M * g.D_in) return;
  const int m = idx / g.D_in, d = idx - m * g.D_in;
  const bool dimwise = a.variant == GPODE_RBF_DIMWISE;
  const float z = a.Z[idx];
  float dz = 0.f;
  for (int k = 0; k < g.D_out; ++k) {
    const float e = dimwise ? a.ell[k * g.D_in + d] : a.ell[d];
    const float var_k = dimwise ? a.var[k] : a.var[0];
    const float inv = 1.f / (e * e);
    float zk = 0.f;
    for (int l = 0; l < g.L; ++l) {
      const float nu = dimwise ? a.nu[(static_cast<size_t>(l) * g.D_out + k) * g.M + m] : a.nu[(static_cast<size_t>(l) * g.M + m) * g.D_out + k];
      const size_t base = (static_cast<size_t>(l) * g.D_out + k) * (2 * g.MP2) + m;
      const float dzk = var_k * nu * inv * (a.acc.pg[base * g.DP + d] - z * a.acc.dnu[base]);
      dz += dzk;
      zk = fmaf(z, dzk, zk);
    }
    atomicAdd(&a.acc.dell_z[k * g.DP + d], zk);
  }
  if (a.d_Z) a.d_Z[idx] = dz;
}

__global__ void k_rbf_finalize_rest(const RbfFinalizeArgs a) {
  const RbfGeom& g = a.g;
  const bool dimwise = a.variant == GPODE_RBF_DIMWISE;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  if (a.d_nu) {
    const long total = static_cast<long>(g.L) * g.D_out * g.M;
    for (long i = tid; i < total; i += nthreads) {
      const int m = static_cast<int>(i % g.M);
      const int k = static_cast<int>((i / g.M) % g.D_out);
      const int l = static_cast<int>(i / (static_cast<long>(g.M) * g.D_out));
      const float var_k = dimwise ? a.var[k] : a.var[0];
      const float v = var_k * a.acc.dnu[(static_cast<size_t>(l) * g.D_out + k) * (2 * g.MP2) + m];
      if (dimwise) a.d_nu[i] = v;
      else a.d_nu[(static_cast<size_t>(l) * g.M + m) * g.D_out + k] = v;
    }
  }
  if (a.d_ell) {
    if (dimwise) {
      for (int i = tid; i < g.D_out * g.D_in; i += nthreads) {
        const int k = i / g.D_in, d = i - k * g.D_in;
        a.d_ell[i] = -(a.acc.dell_x[k * g.DP + d] + a.acc.dell_z[k * g.DP + d]) / a.ell[i];
      }
    } else {
      for (int d = tid; d < g.D_in; d += nthreads) {
        float v = 0.f;
        for (int k = 0; k < g.D_out; ++k) v += a.acc.dell_x[k * g.DP + d] + a.acc.dell_z[k * g.DP + d];
        a.d_ell[d] = -v / a.ell[d];
      }
    }
  }
  if (a.d_var) {
    if (dimwise) {
      for (int k = tid; k < g.D_out; k += nthreads) a.d_var[k] = a.acc.dvar[k] / a.var[k];
    } else if (tid == 0) {
      float v = 0.f;
      for (int k = 0; k < g.D_out; ++k) v += a.acc.dvar[k];
      a.d_var[0] = v / a.var[0];
    }
  }
}

cudaError_t rbf_launch_finalize(const RbfFinalizeArgs& a, cudaStream_t st) {
  const int n = a.g.M * a.g.D_in;
  k_rbf_finalize_z<<<(n + 127) / 128, 128, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  k_rbf_finalize_rest<<<32, 256, 0, st>>>(a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// DP dispatch
// ---------------------------------------------------------------------------------------------
#define GPODE_DP_SWITCH(fn, a, st)            \
  switch ((a).g.DP) {                         \
    case 2: return fn<2>(a, st);              \
    case 4: return fn<4>(a, st);              \
    case 6: return fn<6>(a, st);              \
    case 8: return fn<8>(a, st);              \
    case 10: return fn<10>(a, st);            \
    case 12: return fn<12>(a, st);            \
    case 14: return fn<14>(a, st);            \
    case 16: return fn<16>(a, st);            \
    default: return cudaErrorInvalidValue;    \
  }

cudaError_t rbf_launch_field_fwd(const RbfFieldFwdArgs& a, cudaStream_t st) { GPODE_DP_SWITCH(rbf_field_fwd_dp, a, st) }
cudaError_t rbf_launch_field_bwd(const RbfFieldBwdArgs& a, cudaStream_t st) { GPODE_DP_SWITCH(rbf_field_bwd_dp, a, st) }
cudaError_t rbf_launch_rollout_fwd(const RbfRolloutFwdArgs& a, cudaStream_t st) { GPODE_DP_SWITCH(rbf_rollout_fwd_dp, a, st) }
cudaError_t rbf_launch_rollout_bwd(const RbfRolloutBwdArgs& a, cudaStream_t st) { GPODE_DP_SWITCH(rbf_rollout_bwd_dp, a, st) }
cudaError_t rbf_launch_pgrad(const RbfPgradArgs& a, cudaStream_t st) { GPODE_DP_SWITCH(rbf_pgrad_dp, a, st) }

}  // namespace gpode
