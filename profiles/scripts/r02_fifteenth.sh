#!/bin/bash
# r02q: one shared-memory carve-out preference for every kernel of the iteration (concurrent streams)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { # name workload extra...
  name=$1; wl=$2; shift 2
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02q_${name}.json 2> gpurun_out/r02q_${name}.err
  python - gpurun_out/r02q_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f late %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"], ph.get("late_expansion", 0)))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-500:])
PY
}
run c100 quadrotor
TO_CARVEOUT=-1 run cdefault quadrotor
TO_CARVEOUT=50 run c50 quadrotor
TO_CARVEOUT=0 run c0 quadrotor
TO_NO_OVERLAP=1 run c100_nooverlap quadrotor
TO_SIDE_PRIORITY=0 run c100_noprio quadrotor
run c100_fullstate quadrotor_fullstate
TO_CARVEOUT=-1 run cdefault_fullstate quadrotor_fullstate
