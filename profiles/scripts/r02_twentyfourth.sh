#!/bin/bash
# r02zz: the compact late list + half-warp late pass as defaults: GPU tests, every workload, combinations
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=$PWD/trajectoryoptimization.jl_b200/variants
run() { # name lib workload extra...
  name=$1; lib=$2; wl=$3; shift 3
  LIBTRAJOPT_B200=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02zz_${name}.json 2> gpurun_out/r02zz_${name}.err
  python - gpurun_out/r02zz_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f late %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"], ph.get("late_expansion", 0)))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
}
timeout 600 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02zz_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02zz_tests.log; tail -3 gpurun_out/r02zz_tests.log
run base "" quadrotor
run el128 $V/lib_el128.so quadrotor
run fr6 $V/lib_fr6.so quadrotor
for wl in quadrotor_calm quadrotor_fullstate cartpole acrobot quadrotor_lie; do
  run new_$wl "" $wl
  TO_LATE_LIST=0 TO_FWD_LANES_P2=32 run old_$wl "" $wl
done
TO_LATE_LIST=0 TO_FWD_LANES_P2=32 run old_quadrotor "" quadrotor
