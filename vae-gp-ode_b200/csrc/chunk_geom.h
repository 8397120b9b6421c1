// Geometry of the streamed parameter rows (shared by host code and kernels).
//
// The rows of one sample are a sequence of K blocks; block k holds SP2 "feature" rows of rowf_s floats
// followed (tail == 0, RBF) by MP2 "inducing" rows of rowf_m floats.  With tail == 1 (DF) the blocks
// hold feature rows only and ONE inducing section follows the K blocks.  Rows stream through shared
// memory in chunks of <= RCs / RCm rows that never straddle a section; one field evaluation consumes
// the whole sequence once, in order.
#pragma once
namespace gpode {
struct ChunkGeom {   // constant per launch (lives in the kernel parameter bank, not in registers)
  int stage_floats;  // floats of one pipeline stage (largest chunk)
  int rowf_s, rowf_m;
  int SP2, MP2, NCs, NCm, RCs, RCm;
  int K;             // number of blocks
  int blk_floats;    // floats per block
  int tail;          // 0: inducing section inside every block; 1: once after the K blocks
};
// chunks consumed by one field evaluation
inline __host__ __device__ int chunks_per_eval(const ChunkGeom& cg) { return cg.tail ? cg.K * cg.NCs + cg.NCm : cg.K * (cg.NCs + cg.NCm); }
}  // namespace gpode
