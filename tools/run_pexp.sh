#!/bin/bash
cd "$(dirname "$0")/.."
cp vae-gp-ode_b200/libgpode.so /tmp/libgpode_keep.so
for f in vae-gp-ode_b200/exp/libgpode_pexp*.so; do
  cp "$f" vae-gp-ode_b200/libgpode.so
  echo "== $f"; python tools/dbg_bwd_tc.py 2 2>&1 | grep "bwd tc"
done
cp /tmp/libgpode_keep.so vae-gp-ode_b200/libgpode.so
