#!/usr/bin/env python3
"""Attribute ncu stall samples / executed instructions of one kernel to CUDA source lines.
usage: ncu_lines.py report.ncu-rep cubin kernel_substring [top]
Joins the SASS source page of the report with `nvdisasm -g` line info by instruction order."""
import collections, csv, io, re, subprocess, sys
rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
# split per function
cur, line, insts, infn = None, None, [], False
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        infn = kname in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m: line = (m.group(1), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: insts.append((int(m.group(1), 16), line, m.group(2)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Address", "Kernel Name"): break
    data.append(r)
assert len(data) == len(insts), (len(data), len(insts))
samp = collections.Counter(); ins = collections.Counter()
for r, (off, ln, txt) in zip(data, insts):
    samp[ln] += int(r[idx['# Samples']]); ins[ln] += int(r[idx['Instructions Executed']])
tot = sum(samp.values()); toti = sum(ins.values())
print(f"total samples {tot}, warp-instructions {toti}")
srcs = {}
for (f, n), c in samp.most_common(top):
    if f not in srcs:
        try: srcs[f] = open(f"/root/repo/trajectoryoptimization.jl_b200/csrc/{f}").read().splitlines()
        except Exception: srcs[f] = []
    text = srcs[f][n - 1].strip()[:110] if 0 < n <= len(srcs[f]) else ""
    print(f"{100*c/tot:5.1f}% samp {100*ins[(f,n)]/toti:5.1f}% inst  {f}:{n}  {text}")
