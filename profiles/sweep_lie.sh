#!/bin/bash
# run the error-state bench for every lie.cu variant built by build_lie_variants.sh (on the GPU box)
cd "$(dirname "$0")/.."
echo "default: $(python profiles/bench_phases.py --workload quadrotor_lie --steps 10 --warmup 3 --no-e2e)"
for so in trajectoryoptimization.jl_b200/variants/lib_lie_*.so; do
  echo "$(basename $so .so): $(LIBTRAJOPT_B200=$PWD/$so python profiles/bench_phases.py --workload quadrotor_lie --steps 10 --warmup 3 --no-e2e)"
done
