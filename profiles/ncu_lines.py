#!/usr/bin/env python3
"""Per-source-line hot spots of one kernel: joins the SASS view of an .ncu-rep (samples / executed per instruction)
with `nvdisasm -g` line info of the same cubin (matched by instruction order).
usage: ncu_lines.py report.ncu-rep object.o|lib.so kernel_substring [block_index] [units]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj, kern = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
units = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][which]
hdr = rows[start]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Address", "Kernel Name"): break
    data.append(r)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
lines = []
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines()
        inside = False; cur = None
        for l in out:
            if l.startswith("//---") and ".text." in l: inside = kern in l; cur = None; continue
            if not inside: continue
            m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
            if m: cur = (m.group(1), int(m.group(2))); continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
        if lines: break
print(f"SASS instructions: ncu {len(data)}  nvdisasm {len(lines)}")
n = min(len(data), len(lines))
tot = sum(int(r[idx['# Samples']]) for r in data)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for i in range(n):
    a = agg[lines[i]]
    a[0] += int(data[i][idx['# Samples']]); a[1] += int(data[i][idx['Instructions Executed']])
    for h in stalls: a[2][h[6:]] += int(data[i][idx[h]])
srcfile = {}
print(f"total samples {tot}")
for key, (s, ex, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    text = ""
    if key:
        path = os.path.join(os.path.dirname(obj), "..", key[0]) if not os.path.exists(key[0]) else key[0]
        for cand in (path, os.path.join(os.path.dirname(obj), key[0]), os.path.join(os.path.dirname(os.path.dirname(obj)), "csrc", key[0])):
            if os.path.exists(cand):
                srcfile.setdefault(cand, open(cand).read().splitlines()); text = srcfile[cand][key[1] - 1].strip()[:90]; break
    print(f"{100 * s / tot:5.1f}%  {ex / units:7.1f} inst  {key}  {[(k, round(100 * v / max(s, 1))) for k, v in st.most_common(2)]}  {text}")
