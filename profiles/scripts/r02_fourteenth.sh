#!/bin/bash
# r02p: cost expansion of the early-accepted instances first on the main stream
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { # name workload extra...
  name=$1; wl=$2; shift 2
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02p_${name}.json 2> gpurun_out/r02p_${name}.err
  python - gpurun_out/r02p_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g e2e %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f late %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"], ph.get("late_expansion", 0)))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-500:])
PY
}
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02p_tests.log
tail -4 gpurun_out/r02p_tests.log
run quadrotor quadrotor
run calm quadrotor_calm
run quadrotor2 quadrotor
