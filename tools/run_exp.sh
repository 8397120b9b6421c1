#!/bin/bash
# timing experiments: runs the tc-vs-mma parity check and the T=3 bench against each library variant in vae-gp-ode_b200/exp/
cd "$(dirname "$0")/.."
cp vae-gp-ode_b200/libgpode.so /tmp/libgpode_keep.so
for f in vae-gp-ode_b200/exp/libgpode_*.so; do
  cp "$f" vae-gp-ode_b200/libgpode.so
  timeout 120 python tools/dbg_bwd_tc.py 2 2>&1 | grep "shape\|rror" | head -2
  timeout 120 python bench.py --workload cfg5_t3 --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$f', 'bwd_call_ms', d['roofline']['bwd_call_ms'], 'fwd', d['roofline']['fwd_call_ms'])"
done
cp /tmp/libgpode_keep.so vae-gp-ode_b200/libgpode.so
