#!/bin/bash
# round-2 FINAL artefact run (second edition, after the blocked / staged record cost expansion) on one B200: GPU tests, smoke, the bench lines
# (value + e2e + cpu_baseline for the BASELINE workload, value + e2e for the others), the CPU arm, the launch list, the ncu --set full captures
# of the hot kernels and the compute-sanitizer passes.  Results go to gpurun_out/r02z_*; the ones that are evidence are copied to profiles/.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nproc > gpurun_out/r02z_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r02z_host.txt 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02z_smoke.log; tail -3 gpurun_out/r02z_smoke.log
timeout 900 python -m pytest tests -m gpu -q --tb=line > gpurun_out/r02z_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02z_tests.log; tail -4 gpurun_out/r02z_tests.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d.get("roofline") or {}; ph=r.get("phase_ms") or {}
    print("%-44s value %.4g ms/step %.4g e2e %s frac %s cpu %s phases %s" % (sys.argv[1], d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), r.get("frac"), (d.get("cpu_baseline") or {}).get("value"), {k: round(v, 3) for k, v in ph.items()}))
except Exception as e: print("parse failed", sys.argv[1], e)
PY
}
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02z_bench_quadrotor.json 2> gpurun_out/r02z_bench_quadrotor.err; show gpurun_out/r02z_bench_quadrotor.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02z_bench_reference.json 2> gpurun_out/r02z_bench_reference.err; show gpurun_out/r02z_bench_reference.json
# the restart-free variant of the workload with the 72-register build of the Riccati kernel (one wave of first sweeps: its best case)
TO_FRAG_MINB=7 timeout 600 python bench.py --workload quadrotor_calm --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_bench_quadrotor_calm72.json 2> gpurun_out/r02z_bench_quadrotor_calm72.err; show gpurun_out/r02z_bench_quadrotor_calm72.json
TO_FRAG_MINB=7 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02z_bench_quadrotor72.json 2> gpurun_out/r02z_bench_quadrotor72.err; show gpurun_out/r02z_bench_quadrotor72.json
for wl in quadrotor_calm quadrotor_fullstate cartpole acrobot quadrotor_lie; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_bench_$wl.json 2> gpurun_out/r02z_bench_$wl.err; show gpurun_out/r02z_bench_$wl.json
done
# launch list of the default bench command (cold-cache, serialised by ncu: compare shares)
TO_NO_OVERLAP=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02z_launches.log 2>&1
# full captures: the Riccati kernel on the BASELINE inputs and on the restart-free variant, the dynamics expansion, the record expansion, the line search
cap() { # name regex workload skip count
  TO_NO_OVERLAP=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c $5 -o gpurun_out/r02z_$1 python bench.py --workload $3 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02z_ncu_$1.log 2>&1
}
cap costexp k_expansion_rec16 quadrotor 3 1
cap frag k_riccati_frag quadrotor 3 1
TO_FRAG_MINB=7 cap frag_calm k_riccati_frag quadrotor_calm 3 1
cap expand k_expand_lie quadrotor 3 1
cap fwd k_linesearch quadrotor 6 2
# compute-sanitizer over smoke(): racecheck (shared-memory hazards: the staged cost expansion is new), memcheck, synccheck (mbarrier / warp-sync misuse)
for tool in racecheck memcheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_sanitizer_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/r02z_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK|smoke ok|rc=" gpurun_out/r02z_sanitizer_$tool.log | tail -4
done
ls -la gpurun_out/r02z_* | awk '{print $5, $9}' | tail -40
