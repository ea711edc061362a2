#!/bin/bash
# r02y: balance of the two overlapped chains (expansion kernels on the main stream, late trials on the side stream): compact late list with
# half-warp CTAs of the late pass, CTA size of the dynamics expansion, 80-register build of the Riccati kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=$PWD/trajectoryoptimization.jl_b200/variants
run() { # name lib workload extra...
  name=$1; lib=$2; wl=$3; shift 3
  LIBTRAJOPT_B200=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02y_${name}.json 2> gpurun_out/r02y_${name}.err
  python - gpurun_out/r02y_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f late %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"], ph.get("late_expansion", 0)))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
}
run base "" quadrotor
TO_FWD_LANES_P2=16 run p2h "" quadrotor
TO_LATE_LIST=1 TO_FWD_LANES_P2=16 run list_p2h "" quadrotor
TO_LATE_LIST=1 TO_FWD_LANES_P2=16 TO_PARTITION=32 run list_p2h_part32 "" quadrotor
run el128 $V/lib_el128.so quadrotor
TO_LATE_LIST=1 run list_el128 $V/lib_el128.so quadrotor
run fr6 $V/lib_fr6.so quadrotor
run cx10 $V/lib_cx8k.so quadrotor
TO_FWD_LANES_P1=32 run p1full "" quadrotor
