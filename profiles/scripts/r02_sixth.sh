#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=trajectoryoptimization.jl_b200/variants
run() { # name lib workload
  LIBTRAJOPT_B200=$2 timeout 300 python bench.py --workload $3 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02f_$1_$3.json 2> gpurun_out/r02f_$1_$3.err
  python - gpurun_out/r02f_$1_$3.json $1 $3 <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print("%-8s %-16s value %.4g ms/step %.4g frac %.3f E %.3f R %.3f F %.3f L %.3f" % (sys.argv[2], sys.argv[3], d["value"], d["ms_per_step"], r["frac"], r["phase_ms"]["expand"], r["phase_ms"]["backward"], r["phase_ms"]["forward"], r["phase_ms"]["ladder"]))
except Exception as e: print("parse failed", sys.argv[1], e)
PY
}
for wl in quadrotor_calm quadrotor; do
  run base "" $wl
  for v in e4 r_st3 r_w7 r_w2 r_b6; do run $v $PWD/$V/lib_$v.so $wl; done
done
timeout 1500 python -m pytest tests -m gpu -q --tb=line > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02f_tests.log
tail -8 gpurun_out/r02f_tests.log
