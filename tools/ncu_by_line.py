#!/usr/bin/env python3
"""Join an ncu source-page CSV (SASS view) with nvdisasm line info of the in-tree .so:
executed warp-instructions and stall samples per source line / per region.
usage: ncu -i X.ncu-rep --page source --csv | tools/ncu_by_line.py <kernel-substr-in-mangled-name> <kernel-substr-in-ncu-name> [top]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "aas_enhancement_b200", "libaas_lmfb.so")
mangled, ncuname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], stdout=subprocess.PIPE,
                         text=True).stdout
lines, fn, cur = [], None, "?"
for line in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        fn = m.group(1); continue
    if line.strip().startswith(".section"):
        fn = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = os.path.basename(m.group(1)) + ":" + m.group(2); continue
    if fn and mangled in fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line) and ".byte" not in line and ".dword" not in line:
        lines.append(cur)
rows = list(csv.reader(sys.stdin))
kern, hdr, insts = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        kern = r[1]; hdr = None
    elif r[0] == "Address":
        hdr = r
    elif hdr is not None and kern and ncuname in kern:
        insts.append(dict(zip(hdr, r)))
print(f"sass instrs: nvdisasm {len(lines)}  ncu {len(insts)}")
n = min(len(lines), len(insts))
ex = collections.Counter(); sm = collections.Counter()
for k in range(n):
    ex[lines[k]] += int(insts[k]["Instructions Executed"] or 0)
    sm[lines[k]] += int(insts[k]["# Samples"] or 0)
tot_ex, tot_sm = sum(ex.values()), sum(sm.values())
def _core_ranges():
    path = os.path.join(ROOT, "aas_enhancement_b200", "csrc", "lmfb_core.cuh")
    starts = []
    for i, line in enumerate(open(path), 1):
        m = re.match(r"LMFB_HD\s+\S+\s+(\w+)\(", line)
        if m:
            starts.append((i, m.group(1)))
    return starts
_RANGES = _core_ranges()
def region(key):
    f, l = key.split(":"); l = int(l)
    if f == "fft_codelets.cuh": return "fft32"
    if f == "lmfb_core.cuh":
        name = "core:?"
        for s0, nm in _RANGES:
            if l >= s0 - 1: name = "core:" + nm
        return name
    return f
reg_ex = collections.Counter(); reg_sm = collections.Counter()
for k in ex:
    reg_ex[region(k)] += ex[k]; reg_sm[region(k)] += sm[k]
print("by file: executed%  samples%")
for k, v in reg_ex.most_common():
    print(f"  {k:28s} {100*v/tot_ex:5.1f}%  {100*reg_sm[k]/max(tot_sm,1):5.1f}%")
print("top lines by executed warp-instr (exec%, samples%)")
for k, v in ex.most_common(top):
    print(f"  {k:28s} {100*v/tot_ex:5.1f}%  {100*sm[k]/max(tot_sm,1):5.1f}%")
