// Argument blocks of the solver sweep kernels, generic over the field variant (geometry + accumulators).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace gpode {

// Save-buffer indexing: entry (te, c, s) with te = t * stages + i (t = 0 when nothing is kept),
// c a component, s = l * N + n the global state index -> ((te * C + c) * NL + s): coalesced over states.
// Every geometry G exposes: L, N, NL, D_in, D_out, order, off.

template <class G>
struct FieldFwdArgsT {
  G g;
  const float* packed;
  const float* x;    // (L,N,D_in)
  float* f;          // (L,N,D_out)
  float* f_prior;    // (L,N,D_out) or null
};

template <class G>
struct RolloutFwdArgsT {
  G g;
  const float* packed;
  const float* z0;
  int z0_per_sample;
  const float* ts;
  int T, method, keep;   // keep = 1: saves hold every step (backward follows); 0: one step of scratch
  float* traj;           // (L,N,T,D_in)
  float* xsave;          // stage inputs      [(T-1)*stages][D_in ][NL]
  float* ksave;          // stage derivatives [(T-1)*stages][D_in ][NL]
  float* fpsave;         // prior part        [(T-1)*stages][D_out][NL]
};

template <class G, class Acc>
struct RolloutBwdArgsT {
  G g;
  const float* packed;
  const float* ts;
  int T, method;
  const float* xsave;
  const float* ksave;
  const float* fpsave;
  const float* dtraj;    // (L,N,T,D_in)
  float* dz0;            // (L,N,D_in)
  float* gsave;          // stage adjoints (f part) [(T-1)*stages][D_out][NL]
  float* ybar;           // scratch [D_in][NL]
  float* ystage;         // scratch [stages][D_in][NL]
  float* kbar;           // scratch [D_in][NL]
  Acc acc;
};

template <class G, class Acc>
struct FieldBwdArgsT {
  G g;
  const float* packed;
  const float* x;        // (L,N,D_in)
  const float* gout;     // (L,N,D_out)
  const float* f;        // (L,N,D_out)
  const float* f_prior;  // (L,N,D_out)
  float* dx;             // (L,N,D_in)
  float* xsave;          // [D_in][NL]   transposed copies for the parameter-gradient kernels
  float* gsave;          // [D_out][NL]
  Acc acc;
};

}  // namespace gpode
