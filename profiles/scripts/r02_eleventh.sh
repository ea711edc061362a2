#!/bin/bash
# r02l: GPU tests (incl. hybrid models), the persistent cost-expansion kernel, A/B of the regularisation-ladder speculation rounds
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=trajectoryoptimization.jl_b200/variants
run() { # name lib workload extra...
  name=$1; lib=$2; wl=$3; shift 3
  LIBTRAJOPT_B200=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02l_${name}.json 2> gpurun_out/r02l_${name}.err
  python - gpurun_out/r02l_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g e2e %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"]))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-500:])
PY
}
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02l_tests.log
tail -15 gpurun_out/r02l_tests.log
run base "" quadrotor
run base_calm "" quadrotor_calm
for v in 1 2 3; do
  run rounds$v $PWD/$V/lib_rounds$v.so quadrotor
  LIBTRAJOPT_B200=$PWD/$V/lib_rounds$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -k "fragment or restart or error_state" 2>&1 | tail -2
done
run base2 "" quadrotor
