// RBF sweep policy for SMALL batches (the shape the reference actually trains at: 25 .. 256 trajectories, BASELINE config 1 / 3).
//
// With a few dozen states the regular kernels are latency bound: one warp owns 32 states and walks all D_out (S + M) parameter rows
// serially (~100 us per field evaluation forward + backward at config 1).  Here a CTA still owns 32 states (lane <-> state) but
// brings 16 warps: the parameter rows of the sample are loaded into shared memory ONCE per launch (config 1: 68 KB -- no streaming
// pipeline, no barrier per chunk), every warp evaluates an interleaved 1/16 of the rows of output k for all 32 states, and the
// partial sums meet through a double-buffered shared-memory exchange (one CTA barrier per output dimension).  Warp 0 -- the state
// threads of the generic solver glue (sweep.cuh) -- finishes each output: prior / update parts in the forward, dx_k and the
// lengthscale / variance statistics in the reverse sweep.  Arithmetic per row is the FFMA family's (row_dot / row_bwd_one).
//
// One SM issues at most 4 warp instructions per cycle, and one evaluation at config 1 is ~33,000 of them (measured: 15,000 cycles per
// forward evaluation, instruction-throughput bound).  When the launch leaves SMs idle the OUTPUT dimensions are therefore split over a
// thread-block cluster (grid z = cluster rank, up to 8 CTAs): CTA r evaluates outputs r, r + C, ..; the results are all-gathered through
// distributed shared memory (st.shared::cluster into every CTA of the cluster) and one cluster barrier per evaluation; every CTA then
// runs the (tiny) solver glue redundantly on identical values, so the generic sweep kernels need no cross-CTA logic at all.  In the
// reverse sweep the partial J^T g of every CTA is all-gathered the same way; statistics of output k are accumulated only by its owner.
#pragma once

#include "rbf_kernels.cuh"

namespace gpode {


__device__ __forceinline__ void cluster_sync_all() {   // release / acquire at cluster scope: orders shared::cluster and global accesses
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t local_saddr, int rank, float v) {   // store into CTA `rank`'s copy of a shared-memory location
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

struct SmallSmem {
  float* rows;   // [D_out][SP2 + MP2][ROWF] of this sample
  float* hdr;    // [D_out][hdr_floats]
  float* xs;     // [DP][32]
  float* dx;     // [DP][32]
  float* red;    // [2][kSmWarps][DP + 2][32] partial sums
  float* dell;   // [D_out][DP], then dvar [D_out]
  float* dvar;
  float* gfp;    // [D_out][32] upstream gradient of the evaluation in flight (reverse sweep)
  float* outb;   // [2][D_out][2][32] all-gathered (f_prior, f_update) of one evaluation, double-buffered over evaluations (cluster launches)
  float* dxp;    // [2][kSmMaxCluster][DP][32] all-gathered partial J^T g
  int ev;        // running evaluation counter of this CTA (buffer parity)
};

template <int DP_>
struct RbfSmallPolicy {
  static constexpr int DP = DP_;
  static constexpr int R = 1;
  static constexpr int kThreads = kSmThreads;
  static constexpr int kMinBlocks = 1;
  static constexpr int kStateThreads = kSmStates;
  static constexpr int kXsStride = kSmStates;
  static constexpr bool kCoopGlue = true;   // sweep.cuh: the solver glue (dependent L2 round trips) is spread over the CTA's 512 threads
  static constexpr int kThreadsBwd = kSmThreads;
  static constexpr int kMinBlocksBwd = 1;
  using Geom = RbfGeom;
  using Accum = RbfAccum;
  using Smem = SmallSmem;
  static constexpr int ROW4 = (DP_ + 2) / 2;
  static constexpr int RED = DP_ + 2;

  __device__ static __forceinline__ Smem carve(float* smem, const Geom& g) {
    Smem s;
    s.rows = smem;
    s.hdr = s.rows + g.D_out * (g.SP2 + g.MP2) * g.row_floats;
    s.xs = s.hdr + g.D_out * g.hdr_floats;
    s.dx = s.xs + DP * kSmStates;
    s.red = s.dx + DP * kSmStates;
    s.dell = s.red + 2 * kSmWarps * RED * kSmStates;
    s.dvar = s.dell + g.D_out * DP;
    s.gfp = s.dvar + g.D_out;
    s.outb = s.gfp + g.D_out * kSmStates;
    s.dxp = s.outb + 2 * g.D_out * 2 * kSmStates;
    s.ev = 0;
    return s;
  }
  __device__ static __forceinline__ long setup(Smem& sm, ChunkPipe&, const Geom& g, const float* packed, long, bool) {
    const int l = blockIdx.y;
    const float4* src = reinterpret_cast<const float4*>(rbf_rows_ptr(packed, g, l));
    float4* dst = reinterpret_cast<float4*>(sm.rows);
    const int n4 = g.D_out * (g.SP2 + g.MP2) * g.row_floats / 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
    const float* hdr = rbf_hdr_ptr(packed, g, l);
    for (int i = threadIdx.x; i < g.D_out * g.hdr_floats; i += blockDim.x) sm.hdr[i] = hdr[i];
    for (int i = threadIdx.x; i < 2 * DP * kSmStates; i += blockDim.x) sm.xs[i] = 0.f;   // xs and dx (padded components stay 0)
    for (int i = threadIdx.x; i < g.D_out * (DP + 1); i += blockDim.x) sm.dell[i] = 0.f;  // dell and dvar
    __syncthreads();
    return 0;
  }
  __device__ static __forceinline__ void finish(Smem&) {}
  __device__ static __forceinline__ void flush(const Smem& sm, const Geom& g, const Accum& acc) {
    for (int i = threadIdx.x; i < g.D_out * DP; i += blockDim.x) atomicAdd(&acc.dell_x[i], sm.dell[i]);
    for (int i = threadIdx.x; i < g.D_out; i += blockDim.x) atomicAdd(&acc.dvar[i], sm.dvar[i]);
  }

  template <class Store>
  __device__ static __forceinline__ void eval_fwd(ChunkPipe&, const Geom& g, long, Smem& sm, Store&& store) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = gridDim.z, rank = blockIdx.z;   // cluster = the z extent of the grid (1: no cluster)
    float* outb = sm.outb + (sm.ev & 1) * g.D_out * 2 * kSmStates;
    __syncthreads();   // the stage input staged by warp 0 is visible; the previous evaluation's exchange buffers are free
    float x[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) x[d] = sm.xs[d * kSmStates + lane];
    const int nrows = g.SP2 + g.MP2;
    for (int k = rank; k < g.D_out; k += C) {
      const float* hdr_k = sm.hdr + k * g.hdr_floats;
      float A = 0.f;
#pragma unroll
      for (int d = 0; d < DP; ++d) A = fmaf(hdr_k[d] * x[d], x[d], A);
      const float4* rows = reinterpret_cast<const float4*>(sm.rows + static_cast<size_t>(k) * nrows * g.row_floats);
      float2 accp = make_float2(0.f, 0.f), accu = make_float2(0.f, 0.f);
      // two rows per iteration: independent dot product -> MUFU chains (a warp walks its rows alone; nothing else hides the latency)
      float2 accp2 = make_float2(0.f, 0.f), accu2 = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int j = warp; j < g.SP2; j += 2 * kSmWarps) {
        float4 a[ROW4], b[ROW4];
        const bool two = j + kSmWarps < g.SP2;
        load_row<DP>(a, rows + j * ROW4);
        load_row<DP>(b, rows + (two ? j + kSmWarps : j) * ROW4);
        const float2 wb = two ? hi(b[ROW4 - 1]) : make_float2(0.f, 0.f);
        accp = fma2(cos_2(row_dot<DP, 1>(a, x, 0.f)), hi(a[ROW4 - 1]), accp);
        accp2 = fma2(cos_2(row_dot<DP, 1>(b, x, 0.f)), wb, accp2);
      }
#pragma unroll 1
      for (int j = g.SP2 + warp; j < nrows; j += 2 * kSmWarps) {
        float4 a[ROW4], b[ROW4];
        const bool two = j + kSmWarps < nrows;
        load_row<DP>(a, rows + j * ROW4);
        load_row<DP>(b, rows + (two ? j + kSmWarps : j) * ROW4);
        const float2 wb = two ? hi(b[ROW4 - 1]) : make_float2(0.f, 0.f);
        accu = fma2(ex2_2(row_dot<DP, 1>(a, x, A)), hi(a[ROW4 - 1]), accu);
        accu2 = fma2(ex2_2(row_dot<DP, 1>(b, x, A)), wb, accu2);
      }
      accp.x += accp2.x;
      accp.y += accp2.y;
      accu.x += accu2.x;
      accu.y += accu2.y;
      const int buf = (k / C) & 1;
      float* red = sm.red + (buf * kSmWarps + warp) * RED * kSmStates;
      red[lane] = accp.x + accp.y;
      red[kSmStates + lane] = accu.x + accu.y;
      __syncthreads();
      if (warp == 0) {
        float fp[1] = {0.f}, fu[1] = {0.f};
        const float* all = sm.red + buf * kSmWarps * RED * kSmStates;
#pragma unroll
        for (int w = 0; w < kSmWarps; ++w) {
          fp[0] += all[w * RED * kSmStates + lane];
          fu[0] += all[w * RED * kSmStates + kSmStates + lane];
        }
        fu[0] *= kInvLn2;   // inducing-row weights carry ln2
        if (C == 1) {
          store(k, fp, fu);
        } else {   // all-gather: output k of this evaluation goes to every CTA of the cluster
          const uint32_t a0 = smem_u32(outb + (2 * k) * kSmStates + lane);
          for (int r = 0; r < C; ++r) {
            st_cluster_f32(a0, r, fp[0]);
            st_cluster_f32(a0 + kSmStates * 4, r, fu[0]);
          }
        }
      }
    }
    if (C > 1) {
      cluster_sync_all();
      if (warp == 0)
        for (int k = 0; k < g.D_out; ++k) {
          const float fp[1] = {outb[(2 * k) * kSmStates + lane]}, fu[1] = {outb[(2 * k + 1) * kSmStates + lane]};
          store(k, fp, fu);
        }
    }
    ++sm.ev;
  }

  __device__ static __forceinline__ void vjp(ChunkPipe&, const Geom& g, long, Smem& sm, const States<R>&, const float* gvec, const float* fvec,
                                             const float* fpvec, long kstride, long sstride) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = gridDim.z, rank = blockIdx.z;   // cluster = the z extent of the grid (1: no cluster)
    // upstream gradient and saved outputs of every output dimension go to shared memory in ONE round trip (element (k, state) <-> thread);
    // read one output at a time by warp 0 they cost a dependent L2 latency per output
    for (int e = threadIdx.x; e < g.D_out * kSmStates; e += blockDim.x) {
      const int k = e / kSmStates, sl = e - k * kSmStates;
      const long n = static_cast<long>(blockIdx.x) * kSmStates + sl;
      const bool ok = n < g.N;
      const long at = k * kstride + (static_cast<long>(blockIdx.y) * g.N + (ok ? n : g.N - 1)) * sstride;
      const float gk = ok ? gvec[at] : 0.f;
      sm.gfp[e] = gk;
      const float v = warp_sum(gk * (fvec[at] - 0.5f * fpvec[at]));   // variance statistic of output k, finished here (a warp's lanes share k)
      if (lane == 0 && k % C == rank) atomicAdd(&sm.dvar[k], v);      // (by the CTA that owns output k)
    }
    __syncthreads();
    float x[1][DP], dxs = 0.f;   // dxs: (warp d < DP) component d of J^T g, summed over the outputs
#pragma unroll
    for (int d = 0; d < DP; ++d) x[0][d] = sm.xs[d * kSmStates + lane];
    const int nrows = g.SP2 + g.MP2;
    for (int k = rank; k < g.D_out; k += C) {
      const float* hdr_k = sm.hdr + k * g.hdr_floats;
      float A = 0.f, Q[DP], Es = 0.f;
#pragma unroll
      for (int d = 0; d < DP; ++d) {
        A = fmaf(hdr_k[d] * x[0][d], x[0][d], A);
        Q[d] = 0.f;
      }
      const float4* rows = reinterpret_cast<const float4*>(sm.rows + static_cast<size_t>(k) * nrows * g.row_floats);
      float Q2[DP], Es2 = 0.f;   // second accumulator set: two independent rows per iteration (see eval_fwd)
#pragma unroll
      for (int d = 0; d < DP; ++d) Q2[d] = 0.f;
#pragma unroll 1
      for (int j = warp; j < g.SP2; j += 2 * kSmWarps) {
        float4 a[ROW4], b[ROW4];
        const bool two = j + kSmWarps < g.SP2;
        load_row<DP>(a, rows + j * ROW4);
        load_row<DP>(b, rows + (two ? j + kSmWarps : j) * ROW4);
        const float2 wb = two ? hi(b[ROW4 - 1]) : make_float2(0.f, 0.f);
        row_bwd_one<DP, 1, false>(a, hi(a[ROW4 - 1]), x[0], kHalfPi, Q, Es);
        row_bwd_one<DP, 1, false>(b, wb, x[0], kHalfPi, Q2, Es2);
      }
#pragma unroll 1
      for (int j = g.SP2 + warp; j < nrows; j += 2 * kSmWarps) {
        float4 a[ROW4], b[ROW4];
        const bool two = j + kSmWarps < nrows;
        load_row<DP>(a, rows + j * ROW4);
        load_row<DP>(b, rows + (two ? j + kSmWarps : j) * ROW4);
        const float2 wb = two ? hi(b[ROW4 - 1]) : make_float2(0.f, 0.f);
        row_bwd_one<DP, 1, true>(a, hi(a[ROW4 - 1]), x[0], A, Q, Es);
        row_bwd_one<DP, 1, true>(b, wb, x[0], A, Q2, Es2);
      }
#pragma unroll
      for (int d = 0; d < DP; ++d) Q[d] += Q2[d];
      Es += Es2;
      const int buf = (k / C) & 1;
      float* red = sm.red + (buf * kSmWarps + warp) * RED * kSmStates;
#pragma unroll
      for (int d = 0; d < DP; ++d) red[d * kSmStates + lane] = Q[d];
      red[DP * kSmStates + lane] = Es;
      __syncthreads();
      // finishing output k is spread over DP warps (measured with one finishing warp: 45 % of all warp time at this barrier, the other
      // 15 warps waiting for it): warp d sums component d and the weighted sum over the 16 partials, forms dx_kd and its lengthscale statistic
      if (warp < DP) {
        const float* all = sm.red + buf * kSmWarps * RED * kSmStates;
        float qs = 0.f, es = 0.f;
#pragma unroll
        for (int w = 0; w < kSmWarps; ++w) {
          qs += all[(w * RED + warp) * kSmStates + lane];
          es += all[(w * RED + DP) * kSmStates + lane];
        }
        const float xd = sm.xs[warp * kSmStates + lane];
        const float dxk = sm.gfp[k * kSmStates + lane] * fmaf(2.f * hdr_k[warp] * xd, es, qs);   // (g = 0 for padded states)
        dxs += dxk;
        const float u = warp_sum(xd * dxk);
        if (lane == 0) atomicAdd(&sm.dell[k * DP + warp], u);
      }
    }
    if (C == 1) {
      if (warp < DP) sm.dx[warp * kSmStates + lane] = dxs;
    } else {   // all-gather the partial J^T g (the outputs this CTA owns) and sum the C parts
      float* dxp = sm.dxp + (sm.ev & 1) * kSmMaxCluster * DP * kSmStates;
      if (warp < DP) {
        const uint32_t a0 = smem_u32(dxp + (rank * DP + warp) * kSmStates + lane);
        for (int r = 0; r < C; ++r) st_cluster_f32(a0, r, dxs);
      }
      cluster_sync_all();
      if (warp < DP) {
        float v = 0.f;
        for (int r = 0; r < C; ++r) v += dxp[(r * DP + warp) * kSmStates + lane];
        sm.dx[warp * kSmStates + lane] = v;
      }
    }
    ++sm.ev;
    __syncthreads();   // dx is consumed by the cooperative glue (all threads)
  }
};

}  // namespace gpode
