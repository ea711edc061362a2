#!/bin/bash
# r02m: gain pool for the speculative regularisation candidates (copy instead of one more sweep), round-boundary A/B on top of it,
# knots-per-group A/B of the record cost-expansion kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=$PWD/trajectoryoptimization.jl_b200/variants
run() { # name lib workload extra...
  name=$1; lib=$2; wl=$3; shift 3
  LIBTRAJOPT_B200=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02m_${name}.json 2> gpurun_out/r02m_${name}.err
  python - gpurun_out/r02m_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"]))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-500:])
PY
}
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02m_tests.log
tail -5 gpurun_out/r02m_tests.log
run base "" quadrotor
run base_calm "" quadrotor_calm
for v in rounds0 rounds1 rounds4; do
  run $v $V/lib_$v.so quadrotor
  LIBTRAJOPT_B200=$V/lib_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -k "fragment or restart or error_state" 2>&1 | tail -1
done
for v in cexp_i1 cexp_i2 cexp_i8 cexp_i4m4 cexp_i2m4; do run $v $V/lib_$v.so quadrotor_calm; done
run base2 "" quadrotor
