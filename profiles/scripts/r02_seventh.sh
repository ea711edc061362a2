#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=trajectoryoptimization.jl_b200/variants
run() { # name lib workload
  LIBTRAJOPT_B200=$2 timeout 300 python bench.py --workload $3 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02g_$1_$3.json 2> gpurun_out/r02g_$1_$3.err
  python - gpurun_out/r02g_$1_$3.json $1 $3 <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-8s %-16s value %.4g ms/step %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f" % (sys.argv[2], sys.argv[3], d["value"], d["ms_per_step"], r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"]))
except Exception as e: print("parse failed", sys.argv[1], e)
PY
}
for wl in quadrotor_calm quadrotor; do
  run base "" $wl
  run e4 $PWD/$V/lib_e4.so $wl
done
timeout 1500 python -m pytest tests -m gpu -q --tb=line -x > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02g_tests.log
tail -6 gpurun_out/r02g_tests.log
TO_NO_OVERLAP=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_expand_lie -s 3 -c 1 -o gpurun_out/r02g_expand python bench.py --workload quadrotor --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02g_ncu.log 2>&1
