#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { # name workload extra...
  name=$1; wl=$2; shift 2
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02i_${name}.json 2> gpurun_out/r02i_${name}.err
  python - gpurun_out/r02i_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"]))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-500:])
PY
}
timeout 600 python -m pytest tests -m gpu -q --tb=line -x -k "riccati or fragment or error_state or full_size" > gpurun_out/r02i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02i_tests.log
tail -6 gpurun_out/r02i_tests.log
for b in 128 4096 16384; do run calm_B$b quadrotor_calm --batch $b; done
run quad_B4096 quadrotor
TO_NO_OVERLAP=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_riccati_frag -s 3 -c 1 -o gpurun_out/r02i_frag_calm python bench.py --workload quadrotor_calm --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02i_ncu.log 2>&1
