#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=line > gpurun_out/r02c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02c_tests.log
tail -30 gpurun_out/r02c_tests.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_memcheck.log 2>&1; tail -5 gpurun_out/r02c_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_racecheck.log 2>&1; tail -5 gpurun_out/r02c_racecheck.log
