#!/usr/bin/env python3
"""thread scaling of the CPU oracle port on this host (for the cpu_baseline / --impl reference numbers)"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import trajopt_b200 as TO
import oracle_binding as OB
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(f, open(f).read().strip())
    except Exception: pass
lib = OB.load_oracle(); lib.orc_set_threads.restype = C.c_int
for thr in (1, 8, 16, 32, 64, 128):
    lib.orc_set_threads(thr)
    prob = TO.problems.quadrotor(B=1024, N=101, cls=OB.OracleProblem)
    TO.rollout(prob); TO.ilqr_step(prob, 1)
    t = time.perf_counter(); TO.ilqr_step(prob, 3); dt = time.perf_counter() - t
    print(f"threads {thr}: {1024*3/dt:.0f} inst-iter/s")
    prob.close()
