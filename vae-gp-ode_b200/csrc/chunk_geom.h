// Geometry of the streamed parameter rows (shared by host code and kernels).
#pragma once
namespace gpode {
struct ChunkGeom {   // constant per launch (lives in the kernel parameter bank, not in registers)
  int stage_floats, row_floats, SP2, MP2, NCs, NCm, RCs, RCm, D_out;
};
}  // namespace gpode
