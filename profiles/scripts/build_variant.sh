#!/bin/bash
# A/B builds: build_variant.sh <name> <source.cu> "<-DFLAG=.. ...>"  ->  trajectoryoptimization.jl_b200/variants/lib_<name>.so
# (git-ignored, travels with gpurun); run with LIBTRAJOPT_B200=<path>.  One source file is recompiled with the flags, the rest is linked as built.
set -e
name=$1; src=$2; flags=$3
cd "$(dirname "$0")/../../trajectoryoptimization.jl_b200/csrc"
make -s > /dev/null
mkdir -p ../variants _build/variants
base=$(basename $src .cu)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC -Xptxas -v $flags \
    -c $src -o _build/variants/${base}_$name.o 2> _build/variants/${base}_$name.log
objs=""
for o in capi rollout sweep riccati riccati_small lie riccati_frag forward; do
  if [ "$o" == "$base" ]; then objs="$objs _build/variants/${base}_$name.o"; else objs="$objs _build/$o.o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o ../variants/lib_$name.so $objs
echo "$name: $(grep -E 'Used|spill' _build/variants/${base}_$name.log | sort | uniq -c | sort -rn | head -3 | tr '\n' ' ')"
