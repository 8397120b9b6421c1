#!/bin/bash
# runs tools/dbg_bwd_tc.py <shape> against the profiling build of the library (per-role cycle counters of the fused tcgen05 reverse sweep)
cd "$(dirname "$0")/.."
cp vae-gp-ode_b200/libgpode.so /tmp/libgpode_keep.so
cp vae-gp-ode_b200/libgpode_prof.so vae-gp-ode_b200/libgpode.so
python tools/dbg_bwd_tc.py "${1:-2}"
cp /tmp/libgpode_keep.so vae-gp-ode_b200/libgpode.so
