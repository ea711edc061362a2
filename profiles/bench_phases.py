#!/usr/bin/env python3
"""run bench.py (forwarding the arguments) and print the step time + per-phase kernel times in one short line"""
import json, os, subprocess, sys
out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "..", "bench.py"), "--no-cpu-baseline"] + sys.argv[1:],
                     capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not line:
    print("bench failed:", out.stderr[-800:]); sys.exit(1)
d = json.loads(line[-1])
ph = {k: round(v, 4) for k, v in d["roofline"]["phase_ms"].items()}
e2e = d["e2e"]["value"] if d.get("e2e") else None
print(f"variant={os.environ.get('TO_RICCATI_VARIANT','-')} ms/step={d['ms_per_step']:.4f} value={d['value']:.0f} e2e={e2e} phases={ph} frac={d['roofline']['frac']:.4f}")
