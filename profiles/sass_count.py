#!/usr/bin/env python3
"""static SASS opcode counts of one kernel of an object file: sass_count.py obj.o kernel_substring"""
import collections, os, re, subprocess, sys, tempfile
obj, kern = os.path.abspath(sys.argv[1]), sys.argv[2]
tmp = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
c = collections.Counter(); inside = False
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    for l in subprocess.run(["nvdisasm", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines():
        if l.startswith("//---") and ".text." in l: inside = kern in l; continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if inside and m: c[m.group(2)] += 1
print("total", sum(c.values()), dict(c.most_common(16)))
