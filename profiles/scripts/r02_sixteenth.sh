#!/bin/bash
# r02r: CTA size of the expansion kernels that run next to the late line-search trials; late-instance list for the later line-search passes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=$PWD/trajectoryoptimization.jl_b200/variants
run() { # name lib workload extra...
  name=$1; lib=$2; wl=$3; shift 3
  LIBTRAJOPT_B200=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02r_${name}.json 2> gpurun_out/r02r_${name}.err
  python - gpurun_out/r02r_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f late %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"], ph.get("late_expansion", 0)))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-500:])
PY
}
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02r_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02r_tests.log
tail -4 gpurun_out/r02r_tests.log
run base "" quadrotor
TO_NO_LATE_LIST=1 run nolist "" quadrotor
run old $V/lib_old128_256.so quadrotor
run et128 $V/lib_et128.so quadrotor
run et32 $V/lib_et32.so quadrotor
run ct64 $V/lib_ct64.so quadrotor
run base_calm "" quadrotor_calm
TO_NO_OVERLAP=1 run base_nooverlap "" quadrotor
