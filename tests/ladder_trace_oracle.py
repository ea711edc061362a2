"""Analysis tool (not collected by pytest, test infrastructure like the oracle it drives): where do the sweeps of the regularisation ladder fail?

Copies oracle/ to a scratch directory, patches `backward_pass` to print one line per sweep (instance, candidate, first knot at which a
pivot of Quu + rho I falls below min(luu) + rho, knot of the failure, rho), runs the BASELINE Quadrotor solve (B = 4096, N = 101, error state,
the arithmetic form of the CUDA kernel) for 23 iterations on the CPU and prints, per iteration: how many instances restart and how far they climb, at which
knot the first sweeps fail, and how well two predictors ("restarted in the previous iteration", "rho_0 <= the highest rho that failed in the
instance's last restart episode") would have announced the restarts.  The numbers quoted in profiles/r02_notes.md section 11 come from this script.
usage: python tests/ladder_trace_oracle.py [iterations]"""
import collections
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_traced_oracle(tmp):
    os.makedirs(os.path.join(tmp, "oracle")); os.makedirs(os.path.join(tmp, "include"))
    for f in ("oracle.hpp", "models.hpp", "oracle_capi.cpp", "Makefile"):
        shutil.copy(os.path.join(ROOT, "oracle", f), os.path.join(tmp, "oracle", f))
    for f in os.listdir(os.path.join(ROOT, "include")):
        shutil.copy(os.path.join(ROOT, "include", f), os.path.join(tmp, "include", f))
    p = os.path.join(tmp, "oracle", "oracle.hpp")
    s = open(p).read()
    pd = "                if (!((a > 0) && (detP > 0) && (s00 > 0) && (detS > 0) && (detP < 1e300) && (detS < 1e300))) { ok = false; break; }"
    assert pd in s, "oracle.hpp changed: update the patch"
    s = s.replace(pd, """                { double cmin = 1e300; for (int i = 0; i < 4; i++) cmin = std::min(cmin, hess[(n + i) * nm + n + i]); cmin += rho;
                  const double th = cmin * 0.999999;
                  if (riskk < 0 && (a < th || detP < th * a || s00 < th || detS < th * s00)) riskk = k; }
""" + pd.replace("{ ok = false; break; }", "{ ok = false; failk = k; break; }"))
    s = s.replace("        bool ok = true;\n        const double rho = P.rho[b];", "        bool ok = true; int failk = -1, riskk = -1;\n        const double rho = P.rho[b];")
    ok = "        if (ok) { P.dV[2 * b] = dV1; P.dV[2 * b + 1] = dV2; break; }"
    assert ok in s
    s = s.replace(ok, '        if (getenv("ORC_TRACE")) {\n#pragma omp critical\n            fprintf(stderr, "SWEEP %d %d %d %d %.3e\\n", b, restarts, riskk, failk, rho); }\n' + ok)
    open(p, "w").write("#include <cstdio>\n#include <cstdlib>\n" + s)
    subprocess.check_call(["make", "-C", os.path.join(tmp, "oracle")], stdout=subprocess.DEVNULL)
    return os.path.join(tmp, "oracle", "_build", "liboracle.so")


def run(lib, iters, trace_path):
    code = f"""
import sys; sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
import oracle_binding as OB
OB.ORACLE_LIB = {lib!r}; OB.ORACLE_DIR = {os.path.dirname(os.path.dirname(lib))!r}
import trajopt_b200 as TO
p = TO.problems.quadrotor(B=4096, N=101, error_state=True, cls=OB.OracleProblem)
p.set_backward_variant(1)
TO.rollout(p)
for it in range({iters}):
    sys.stderr.write("ITER %d\\n" % it); sys.stderr.flush()
    TO.ilqr_step(p, 1)
"""
    with open(trace_path, "w") as f:
        subprocess.check_call([sys.executable, "-c", code], env=dict(os.environ, ORC_TRACE="1"), stderr=f)


def analyse(trace_path):
    it = -1
    sweeps = collections.defaultdict(lambda: collections.defaultdict(list))
    for line in open(trace_path):
        w = line.split()
        if not w: continue
        if w[0] == "ITER": it = int(w[1])
        elif w[0] == "SWEEP": sweeps[it][int(w[1])].append((int(w[2]), int(w[3]), int(w[4]), float(w[5])))   # cand, riskk, failk, rho
    B = 4096
    prev_restarted = np.zeros(B, bool)
    rho_fail = np.full(B, -1.0)
    for it in sorted(sweeps):
        status = np.zeros(B, int); fk0 = []; lead = []
        tp = fp = fn = 0; hist_hit = 0
        newfail = {}
        for b, lst in sweeps[it].items():
            lst.sort()
            status[b] = len(lst) - 1
            cand0 = lst[0]
            failed0 = cand0[2] >= 0
            if failed0:
                fk0.append(cand0[2]); lead.append(cand0[1] - cand0[2])
                hist_hit += int(prev_restarted[b])
            pred = rho_fail[b] >= 0 and cand0[3] <= rho_fail[b]
            tp += int(pred and failed0); fp += int(pred and not failed0); fn += int(failed0 and not pred)
            fails = [r for (_, _, fk, r) in lst if fk >= 0]
            if fails: newfail[b] = max(fails)
            elif pred: rho_fail[b] = -1.0                     # a success at rho_0 <= rho_fail contradicts the prediction: forget the episode
        for b, r in newfail.items(): rho_fail[b] = r
        restarted = status > 0
        fk0 = np.array(fk0) if fk0 else np.array([0])
        print(f"it {it:2d}: restarted {restarted.sum():4d} ({100 * restarted.mean():4.1f} %)  rungs {np.bincount(np.clip(status, 0, 13), minlength=14)[1:].tolist()}  "
              f"first-sweep failure knot pct 10/50/90 {np.percentile(fk0, [10, 50, 90]).round().astype(int).tolist()} (k < 30: {(fk0 < 30).sum()})  "
              f"risk-to-failure lead (knots) median {int(np.median(lead)) if lead else 0}  | predicted by 'restarted last iteration' {hist_hit}/{restarted.sum()}"
              f"  by rho history tp {tp} fp {fp} fn {fn}")
        prev_restarted = restarted


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 23
    with tempfile.TemporaryDirectory() as tmp:
        lib = build_traced_oracle(tmp)
        trace = os.path.join(tmp, "trace.txt")
        run(lib, iters, trace)
        analyse(trace)
