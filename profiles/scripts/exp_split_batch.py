"""Experiment: is one batch of 4096 instances faster as ONE handle or as k handles of 4096/k instances stepping concurrently (each on its own
stream)?  Latency-bound kernels of one sub-batch could overlap the pipe-bound ones of another.  usage: exp_split_batch.py [workload]"""
import ctypes as C, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
import trajopt_b200 as TO
K = TO.capi
wl = sys.argv[1] if len(sys.argv) > 1 else "quadrotor"
N = bench.WORKLOADS[wl]["N"]
for total, k in ((4096, 1), (4096, 2), (4096, 4), (8192, 2), (8192, 1), (16384, 1), (16384, 4)):
    B = total // k
    probs = [bench.build_problem(wl, B, N) for _ in range(k)]
    lib = probs[0]._lib
    for p in probs:
        TO.rollout(p)
    def steps(n):
        for _ in range(n):
            for p in probs:
                K.check(lib, p._h, lib.to_ilqr_step(p._h, 1))
    steps(3); torch.cuda.synchronize()
    t0 = time.perf_counter(); steps(20); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{wl}: {k} handle(s) x {B:5d} instances: {dt / 20 * 1e3:.3f} ms per iteration of all, {total * 20 / dt / 1e6:.3f} M instance-iterations/s", flush=True)
    for p in probs:
        p.close()
