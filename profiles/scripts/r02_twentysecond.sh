#!/bin/bash
# r02x: dynamics expansion with the record's [A_e B_e] block staged in shared memory (k_expand_lie_staged, TO_EXPAND_STAGE=1)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=$PWD/trajectoryoptimization.jl_b200/variants
run() { # name lib workload extra...
  name=$1; lib=$2; wl=$3; shift 3
  LIBTRAJOPT_B200=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02x_${name}.json 2> gpurun_out/r02x_${name}.err
  python - gpurun_out/r02x_${name}.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; ph=r["phase_ms"]
    print("%-22s value %.4g ms/step %.4g frac %.3f E %.3f C %.3f R %.3f F %.3f L %.3f late %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], r["frac"], ph["expand"], ph["cost_expansion"], ph["backward"], ph["forward"], ph["ladder"], ph.get("late_expansion", 0)))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
}
timeout 200 python profiles/scripts/cexp_ab.py gpurun_out/r02x_ab_a.npz > gpurun_out/r02x_ab.log 2>&1
TO_EXPAND_STAGE=1 timeout 200 python profiles/scripts/cexp_ab.py gpurun_out/r02x_ab_b.npz >> gpurun_out/r02x_ab.log 2>&1
python profiles/scripts/cexp_ab.py --cmp gpurun_out/r02x_ab_a.npz gpurun_out/r02x_ab_b.npz >> gpurun_out/r02x_ab.log 2>&1
tail -3 gpurun_out/r02x_ab.log; rm -f gpurun_out/r02x_ab_a.npz gpurun_out/r02x_ab_b.npz
run base "" quadrotor
TO_EXPAND_STAGE=1 run stage "" quadrotor
TO_NO_OVERLAP=1 run seq "" quadrotor
TO_NO_OVERLAP=1 TO_EXPAND_STAGE=1 run seq_stage "" quadrotor
TO_EXPAND_STAGE=1 run stage128 $V/lib_el128.so quadrotor
TO_EXPAND_STAGE=1 TO_LATE_LIST=1 TO_PARTITION=32 run stage_list_part32 "" quadrotor
TO_EXPAND_STAGE=1 timeout 600 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02x_tests.log; tail -3 gpurun_out/r02x_tests.log
