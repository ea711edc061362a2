#!/bin/bash
# r02s: iteration order -- late trials beside the latency-bound cost expansion only, dynamics expansion after the join (TO_ITER_ORDER=1) vs beside it (0)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { # name workload extra...
  name=$1; wl=$2; shift 2
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02s_${name}.json 2> gpurun_out/r02s_${name}.err
  python - gpurun_out/r02s_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f late %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"], ph.get("late_expansion", 0)))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-500:])
PY
}
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02s_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02s_tests.log
tail -4 gpurun_out/r02s_tests.log
for wl in quadrotor quadrotor_calm quadrotor_fullstate cartpole acrobot; do
  run ${wl}_o1 $wl
  TO_ITER_ORDER=0 run ${wl}_o0 $wl
done
TO_ITER_ORDER=0 TO_NO_LATE_LIST=1 run quadrotor_o0_nolist quadrotor
TO_ITER_ORDER=0 TO_NO_LATE_LIST=1 run fullstate_o0_nolist quadrotor_fullstate
