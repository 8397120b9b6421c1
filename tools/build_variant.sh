#!/bin/bash
# builds a variant of libgpode.so into vae-gp-ode_b200/exp/libgpode_<name>.so: only rbf_inst.cu at DP = 16 is recompiled with the extra
# flags, every other object comes from the regular build.   usage: tools/build_variant.sh <name> <extra nvcc flags...>
set -e
cd "$(dirname "$0")/../vae-gp-ode_b200/csrc"
name=$1; shift
mkdir -p ../exp build_var
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include -I. --expt-relaxed-constexpr -diag-suppress 20281 "$@" -DGPODE_DP=16 -c rbf_inst.cu -o build_var/rbf_inst_16_$name.o
objs=$(ls build/*.o | grep -v rbf_inst_16.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../exp/libgpode_$name.so $objs build_var/rbf_inst_16_$name.o -lcudart
echo built ../exp/libgpode_$name.so
