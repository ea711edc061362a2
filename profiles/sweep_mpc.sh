#!/bin/bash
# BASELINE config 5: Quadrotor sweep N in {51,101,201,401} x batch in {256,1024,4096,16384}: step time, throughput and the Riccati
# kernel's HBM-roofline fraction per point (one line each; run on one B200, e.g. under gpurun)
for n in 51 101 201 401; do for b in 256 1024 4096 16384; do
  python - "$n" "$b" <<'PY'
import json, os, subprocess, sys
N, B = sys.argv[1], sys.argv[2]
out = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--no-e2e", "--steps", "10", "--warmup", "3", "--batch", B, "--N", N], capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not line:
    print(f"N={N} B={B} failed: {out.stderr[-300:]}")
else:
    d = json.loads(line[-1]); r = d["roofline"]; ph = r["phase_ms"]
    print(json.dumps({"N": int(N), "B": int(B), "ms_per_step": round(d["ms_per_step"], 4), "inst_iter_per_s": round(d["value"]),
                      "knot_iter_per_s": round(d["value"] * (int(N) - 1)), "riccati_ms": round(ph["backward"], 4), "riccati_GBps": round(r["achieved"], 1),
                      "riccati_frac": round(r["frac"], 4), "expand_ms": round(ph["expand"], 4), "forward_ms": round(ph["forward"], 4), "ladder_ms": round(ph["ladder"], 4)}))
PY
done; done
