#!/bin/bash
# r02u: blocked record cost expansion (k_expansion_rec16b) against k_expansion_rec16; L2 prefetch / ring depth of the Riccati record stream
# (lone-warp sweep latency: quadrotor_calm at 512 instances); whole-ladder speculation.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=$PWD/trajectoryoptimization.jl_b200/variants
run() { # name lib workload extra...
  name=$1; lib=$2; wl=$3; shift 3
  LIBTRAJOPT_B200=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02u_${name}.json 2> gpurun_out/r02u_${name}.err
  python - gpurun_out/r02u_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f late %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"], ph.get("late_expansion", 0)))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-500:])
PY
}
# bit-identity of the two cost-expansion kernels
TO_CEXP_V1=1 timeout 200 python profiles/scripts/cexp_ab.py gpurun_out/r02u_ab_v1.npz > gpurun_out/r02u_ab.log 2>&1
timeout 200 python profiles/scripts/cexp_ab.py gpurun_out/r02u_ab_v2.npz >> gpurun_out/r02u_ab.log 2>&1
python profiles/scripts/cexp_ab.py --cmp gpurun_out/r02u_ab_v1.npz gpurun_out/r02u_ab_v2.npz >> gpurun_out/r02u_ab.log 2>&1
tail -4 gpurun_out/r02u_ab.log
rm -f gpurun_out/r02u_ab_v1.npz gpurun_out/r02u_ab_v2.npz
timeout 600 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02u_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02u_tests.log
tail -3 gpurun_out/r02u_tests.log
run base "" quadrotor
TO_CEXP_V1=1 run cexp_v1 "" quadrotor
TO_CEXP2_UNITS=1 run cexp_u1 "" quadrotor
TO_CEXP2_UNITS=4 run cexp_u4 "" quadrotor
run cx4 $V/lib_cx4.so quadrotor
for v in pf4 pf8 s3 s3pf6 r5; do run $v $V/lib_$v.so quadrotor; done
run lone_base "" quadrotor_calm --batch 512
for v in pf4 pf8 s3 s3pf6; do run lone_$v $V/lib_$v.so quadrotor_calm --batch 512; done
run calm_base "" quadrotor_calm
run calm_pf4 $V/lib_pf4.so quadrotor_calm
