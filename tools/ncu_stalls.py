#!/usr/bin/env python3
"""Summarise an ncu --page source --csv dump: per-kernel stall-reason totals and the hottest
SASS instructions.   usage: ncu -i X.ncu-rep --page source --csv | tools/ncu_stalls.py [topN]"""
import csv
import sys

top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
kern, hdr, data = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        kern = r[1]
        data[kern] = []
        hdr = None
    elif r[0] == "Address":
        hdr = r
    elif hdr is not None and kern is not None:
        data[kern].append(dict(zip(hdr, r)))
for kern, insts in data.items():
    print("=" * 100)
    print(kern, "instructions:", len(insts))
    stall_cols = [c for c in insts[0] if c.startswith("stall_") and "Not Issued" not in c]
    tot = {c: sum(int(i[c] or 0) for i in insts) for c in stall_cols}
    total = sum(tot.values())
    print("total samples", total, " executed warp-instr", sum(int(i["Instructions Executed"] or 0) for i in insts))
    for c, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
        print(f"   {c:28s} {v:8d} {100.0 * v / max(total, 1):5.1f}%")
    print("  hottest instructions (samples, idx, sass, top reasons)")
    order = sorted(range(len(insts)), key=lambda k: -int(insts[k]["# Samples"] or 0))[:top_n]
    for k in order:
        i = insts[k]
        reasons = sorted(((int(i[c] or 0), c) for c in stall_cols), reverse=True)[:2]
        print(f"   {int(i['# Samples']):6d} #{k:5d} {i['Source'].strip()[:70]:70s} " +
              " ".join(f"{c[6:]}={v}" for v, c in reasons if v))
    # coarse phase histogram: samples per 5% of the instruction stream
    nb = 20
    hist = [0] * nb
    for k, i in enumerate(insts):
        hist[min(nb - 1, k * nb // len(insts))] += int(i["# Samples"] or 0)
    print("  samples by position in the kernel (20 buckets):", hist)
