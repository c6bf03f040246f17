#!/usr/bin/env python3
"""Executed warp instructions and stall samples per kernel region (regions split at BAR.SYNC) and the
executed opcode mix, from `ncu -i X.ncu-rep --page source --csv`.
usage: ncu -i X.ncu-rep --page source --csv | tools/ncu_regions.py <tiles per launch>"""
import csv, sys
tiles = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
rows = list(csv.reader(sys.stdin))
kern = None; hdr = None; data = {}
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name": kern = r[1]; data[kern] = []; hdr = None
    elif r[0] == "Address": hdr = r
    elif hdr is not None and kern is not None: data[kern].append(dict(zip(hdr, r)))
for kern, insts in data.items():
    print(kern[:90])
    stall_cols = [c for c in insts[0] if c.startswith("stall_") and "Not Issued" not in c]
    start = 0; acc = 0; samp = 0; st = {}
    tot_e = 0
    for k, i in enumerate(insts):
        ex = int(i["Instructions Executed"] or 0); acc += ex; tot_e += ex; samp += int(i["# Samples"] or 0)
        for c in stall_cols: st[c] = st.get(c, 0) + int(i[c] or 0)
        if "BAR.SYNC" in i["Source"] or k == len(insts) - 1:
            top = " ".join("%s=%d" % (c[6:], v) for c, v in sorted(st.items(), key=lambda kv: -kv[1])[:5])
            print("  region #%5d..%5d  exec/tile %8.0f   samples %6d   %s" % (start, k, acc / tiles, samp, top))
            start = k + 1; acc = 0; samp = 0; st = {}
    mix = {}
    for i in insts:
        t = i["Source"].split()
        op = t[1] if t[0].startswith('@') else t[0]
        op = '.'.join(op.split('.')[:2]) if op.startswith(('LD', 'ST')) else op.split('.')[0]
        mix[op] = mix.get(op, 0) + int(i["Instructions Executed"] or 0) / tiles
    print("  total/tile %.0f   mix/tile:" % (tot_e / tiles), " ".join("%s:%.0f" % (o, v) for o, v in sorted(mix.items(), key=lambda kv: -kv[1])[:26]))
