#!/bin/bash
# r02w: k_expansion_rec16b with its outputs staged in shared memory (whole-line stores); compact late list + SM partition
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=$PWD/trajectoryoptimization.jl_b200/variants
run() { # name lib workload extra...
  name=$1; lib=$2; wl=$3; shift 3
  LIBTRAJOPT_B200=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02w_${name}.json 2> gpurun_out/r02w_${name}.err
  python - gpurun_out/r02w_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f late %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"], ph.get("late_expansion", 0)))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
}
timeout 400 python -m pytest tests/test_gpu_costexp.py -m gpu -q --tb=short -x > gpurun_out/r02w_tests_ab.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02w_tests_ab.log; tail -3 gpurun_out/r02w_tests_ab.log
run base "" quadrotor
TO_NO_OVERLAP=1 run seq "" quadrotor
run cy9 $V/lib_cy9.so quadrotor
run cy96 $V/lib_cy96.so quadrotor
TO_CEXP2_UNITS=2 run u2 "" quadrotor
TO_LATE_LIST=1 run list "" quadrotor
for p in 24 32 40; do TO_LATE_LIST=1 TO_PARTITION=$p run list_part$p "" quadrotor; done
timeout 600 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02w_tests.log; tail -3 gpurun_out/r02w_tests.log
