#!/bin/bash
# r02t: depth of the operand ring of the forward rollout (late trials next to the expansion kernels), iteration order
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=$PWD/trajectoryoptimization.jl_b200/variants
run() { # name lib workload extra...
  name=$1; lib=$2; wl=$3; shift 3
  LIBTRAJOPT_B200=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02t_${name}.json 2> gpurun_out/r02t_${name}.err
  python - gpurun_out/r02t_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f late %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"], ph.get("late_expansion", 0)))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-500:])
PY
}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x > gpurun_out/r02t_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02t_tests.log
tail -3 gpurun_out/r02t_tests.log
for o in 0 1; do
  TO_ITER_ORDER=$o run s4_o$o "" quadrotor
  for ns in 2 3 6; do TO_ITER_ORDER=$o run s${ns}_o$o $V/lib_fs$ns.so quadrotor; done
done
TO_ITER_ORDER=0 TO_NO_LATE_LIST=1 run s4_o0_nolist "" quadrotor
TO_ITER_ORDER=0 run s4_o0_fullstate "" quadrotor_fullstate
TO_ITER_ORDER=1 run s4_o1_fullstate "" quadrotor_fullstate
TO_ITER_ORDER=0 TO_NO_LATE_LIST=1 run s4_o0_nolist_fullstate "" quadrotor_fullstate
