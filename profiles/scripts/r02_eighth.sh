#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
./profiles/microbench/fp64_latency > gpurun_out/r02h_latency.txt 2>&1; cat gpurun_out/r02h_latency.txt
run() { # name workload extra...
  name=$1; wl=$2; shift 2
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02h_${name}.json 2> gpurun_out/r02h_${name}.err
  python - gpurun_out/r02h_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"]))
except Exception as e: print("parse failed", sys.argv[1], e)
PY
}
for b in 128 512 1024 2048 4096 8192 16384; do run calm_B$b quadrotor_calm --batch $b; done
for b in 512 4096; do run quad_B$b quadrotor --batch $b; done
timeout 1500 python -m pytest tests -m gpu -q --tb=line -x > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02h_tests.log
tail -6 gpurun_out/r02h_tests.log
