#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02e_smoke.log; tail -3 gpurun_out/r02e_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --tb=line > gpurun_out/r02e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02e_tests.log
tail -15 gpurun_out/r02e_tests.log
for wl in quadrotor quadrotor_calm quadrotor_fullstate; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_$wl.json 2> gpurun_out/r02e_bench_$wl.err
  python - gpurun_out/r02e_bench_$wl.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.4g ms/step %.4g e2e %.4g frac %.3f phases %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r["phase_ms"]))
except Exception as e: print("parse failed", e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
TO_NO_OVERLAP=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_expand_lie -s 3 -c 1 -o gpurun_out/r02e_expand python bench.py --workload quadrotor --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02e_ncu.log 2>&1
TO_NO_OVERLAP=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_linesearch -s 6 -c 2 -o gpurun_out/r02e_fwd python bench.py --workload quadrotor --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02e_ncu2.log 2>&1
