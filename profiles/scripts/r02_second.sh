#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=line > gpurun_out/r02b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02b_tests.log
tail -25 gpurun_out/r02b_tests.log
timeout 900 python profiles/scripts/parity_budget.py > gpurun_out/r02b_parity_budget.log 2>&1
tail -12 gpurun_out/r02b_parity_budget.log
for wl in quadrotor_calm; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_$wl.json 2> gpurun_out/r02b_bench_$wl.err
  python - gpurun_out/r02b_bench_$wl.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print(sys.argv[1], "value %.4g ms/step %.4g frac %.3f phases %s" % (d["value"], d["ms_per_step"], r["frac"], r["phase_ms"]))
PY
done
TO_NO_OVERLAP=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_riccati_frag -s 3 -c 1 -o gpurun_out/r02b_frag_calm python bench.py --workload quadrotor_calm --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02b_ncu.log 2>&1
