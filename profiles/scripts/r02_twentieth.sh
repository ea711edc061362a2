#!/bin/bash
# r02v: SM partition (green contexts) for the overlapped part of an iteration; register budget of k_expansion_rec16b; sequential phase times
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=$PWD/trajectoryoptimization.jl_b200/variants
run() { # name lib workload extra...
  name=$1; lib=$2; wl=$3; shift 3
  LIBTRAJOPT_B200=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02v_${name}.json 2> gpurun_out/r02v_${name}.err
  python - gpurun_out/r02v_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f late %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"], ph.get("late_expansion", 0)))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
}
timeout 60 profiles/microbench/green_ctx 24 > gpurun_out/r02v_green.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_green.log; cat gpurun_out/r02v_green.log
run base "" quadrotor
for p in 16 24 32 48; do TO_VERBOSE=1 TO_PARTITION=$p run part$p "" quadrotor; grep -h "SM partition" gpurun_out/r02v_part$p.err | head -1; done
TO_NO_OVERLAP=1 run seq "" quadrotor
run cx7 $V/lib_cx7.so quadrotor
run cx8 $V/lib_cx8.so quadrotor
TO_PARTITION=24 run part24_calm "" quadrotor_calm
TO_PARTITION=24 run part24_full "" quadrotor_fullstate
run base_full "" quadrotor_fullstate
# results must not depend on the partition
timeout 200 python profiles/scripts/cexp_ab.py gpurun_out/r02v_ab_a.npz > gpurun_out/r02v_ab.log 2>&1
TO_PARTITION=24 timeout 200 python profiles/scripts/cexp_ab.py gpurun_out/r02v_ab_b.npz >> gpurun_out/r02v_ab.log 2>&1
python profiles/scripts/cexp_ab.py --cmp gpurun_out/r02v_ab_a.npz gpurun_out/r02v_ab_b.npz >> gpurun_out/r02v_ab.log 2>&1
tail -2 gpurun_out/r02v_ab.log; rm -f gpurun_out/r02v_ab_a.npz gpurun_out/r02v_ab_b.npz
if grep -q "GREEN OK" gpurun_out/r02v_green.log; then
  TO_PARTITION=24 timeout 600 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02v_tests_part.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02v_tests_part.log
  tail -3 gpurun_out/r02v_tests_part.log
fi
