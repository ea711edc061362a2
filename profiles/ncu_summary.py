#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page + source page): key metrics, SASS opcode mix, stall reasons.
usage: ncu_summary.py report.ncu-rep [units_per_launch] [launch_index]   (units = warp-knots etc. for per-unit instruction
counts; launch_index = which captured launch the SASS / stall breakdown is for, default 0)"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 and float(sys.argv[2]) > 0 else None
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg']
for li, vals in enumerate(rows[2:]):
    print(f"--- launch {li}: {vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''}")
    for k in keys:
        if k in hdr:
            print(f"  {k} = {vals[hdr.index(k)]} {rows[1][hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# the source page concatenates the launches; take block `which`
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][which]
print(f"=== SASS / stall breakdown of launch {which}")
hdr = rows[start]; idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Address", "Kernel Name"):
        break
    data.append(r)
tot = sum(int(r[idx['# Samples']]) for r in data)
totinst = sum(int(r[idx['Instructions Executed']]) for r in data)
print(f"SASS lines {len(data)}  warp-instructions {totinst}  samples {tot}")
oi = collections.Counter(); os_ = collections.Counter()
for r in data:
    parts = r[idx['Source']].split()
    op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
    oi[op] += int(r[idx['Instructions Executed']]); os_[op] += int(r[idx['# Samples']])
d = units or 1.0
print("instr" + (" per unit" if units else "") + ":", [(k, round(v / d, 1)) for k, v in oi.most_common(24)])
print("samples %:", [(k, round(100 * v / max(tot, 1), 1)) for k, v in os_.most_common(14)])
st = {s: sum(int(r[idx[s]]) for r in data) for s in hdr if s.startswith('stall_') and 'Not Issued' not in s}
print("stalls %:", [(k, round(100 * v / max(tot, 1), 1)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]])
