#!/bin/bash
# round 2, first GPU call: parity of the new fragment Riccati kernel + baseline timings of every workload + one ncu capture
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02a_tests.log
tail -5 gpurun_out/r02a_tests.log
for wl in quadrotor quadrotor_fullstate quadrotor_lie cartpole acrobot; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_$wl.json 2> gpurun_out/r02a_bench_$wl.err
done
TO_NO_FRAG=1 timeout 300 python bench.py --workload quadrotor --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_quadrotor_nofrag.json 2> gpurun_out/r02a_bench_quadrotor_nofrag.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02a_bench_reference.json 2> gpurun_out/r02a_bench_reference.err
TO_NO_OVERLAP=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_riccati_frag -s 3 -c 1 -o gpurun_out/r02a_frag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02a_ncu.log 2>&1
for f in gpurun_out/r02a_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print(" value %.4g ms/step %.4g e2e %s frac %s phases %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), r.get("frac"), r.get("phase_ms")))
except Exception as e: print(" parse failed", e)
PY
done
