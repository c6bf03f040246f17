#!/usr/bin/env python3
"""Copy the evidence of one gpurun round from gpurun_out/ (scratch) into profiles/ (tracked):
bench lines, the ncu launch list + per-kernel shares, selected ncu --set full metrics, stall
summaries, DRAM traffic per launch (profiles/traffic.json) and the SASS of the default kernels.
usage: tools/save_profiles.py <tag>"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1]
os.makedirs(P, exist_ok=True)

for name in (f"bench_{tag}.json", f"bench_sweep_{tag}.json", f"bench_paired_{tag}.json", f"bench_ref_{tag}.json",
             f"modes_{tag}.txt", f"host_{tag}.txt", f"smi_{tag}.txt",
             f"pytest_{tag}.txt", f"smoke_{tag}.txt", f"launches_{tag}.csv"):
    src = os.path.join(G, name)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, f"{tag}_{name.replace('_' + tag, '')}"))

# launch list shares
lpath = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(lpath):
    rows = list(csv.reader(open(lpath)))
    hdr = [r for r in rows if "Kernel Name" in r][0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows:
        if len(r) > vi and r[ki] != "Kernel Name":
            try:
                d[r[ki]].append(float(r[vi].replace(",", "")))
            except ValueError:
                pass
    tot = sum(sum(v) for v in d.values())
    with open(os.path.join(P, f"{tag}_launch_shares.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{sum(v) / tot * 100:6.2f}%  n={len(v):3d}  avg_ns={sum(v) / len(v):12.1f}  {k[:110]}\n")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
traffic = {}
tpath = os.path.join(P, "traffic.json")
if os.path.exists(tpath):
    traffic = json.load(open(tpath))
for wl, rep in (("chime4_30x6s", f"prof_chime_{tag}.ncu-rep"), ("sweep_256x10s", f"prof_sweep_{tag}.ncu-rep")):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(P, f"{tag}_ncu_{wl}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, workload {wl}, kernels matching lmfb_k1 (one launch each)\n")
        seen = set()
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            if name in seen:
                continue
            seen.add(name)
            f.write(f"== {name}\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"   {w:70s} {r[i]:>16s} {units[i]}\n")
            key = "k1_bwd_dram_bytes" if ("<1, 1" in name or "(bool)1" in name) else "k1_fwd_dram_bytes"
            if True:
                def val(m):
                    i = hdr.index(m); x = float(r[i].replace(",", "")); u = units[i].lower()
                    return x * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1)
                traffic.setdefault(wl, {})[key] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
                traffic[wl]["tag"] = tag
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    st = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_stalls.py"), "15"], input=src,
                        stdout=subprocess.PIPE, text=True).stdout
    open(os.path.join(P, f"{tag}_stalls_{wl}.txt"), "w").write(st)
    tiles = {"chime4_30x6s": 30 * 19, "sweep_256x10s": 256 * 32}[wl]
    rg = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), str(tiles)], input=src,
                        stdout=subprocess.PIPE, text=True).stdout
    open(os.path.join(P, f"{tag}_regions_{wl}.txt"), "w").write(
        "# executed warp instructions per tile and stall samples per kernel region (split at BAR.SYNC)\n" + rg)
json.dump(traffic, open(tpath, "w"), indent=1)

# SASS of the library (opcode histogram per kernel + full listing of the default reim kernels)
lib = os.path.join(ROOT, "aas_enhancement_b200", "libaas_lmfb.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
blocks, cur, name = {}, [], None
for line in sass.splitlines():
    if "Function :" in line:
        if name:
            blocks[name] = cur
        name, cur = line.split("Function :")[1].strip(), []
    elif name:
        cur.append(line)
if name:
    blocks[name] = cur
with open(os.path.join(P, f"{tag}_sass_summary.txt"), "w") as f:
    f.write("# cuobjdump -sass aas_enhancement_b200/libaas_lmfb.so: instruction count and top opcodes per kernel\n")
    for k, lines in sorted(blocks.items()):
        ops = collections.Counter()
        for l in lines:
            parts = l.split("*/")
            if len(parts) >= 2 and parts[0].strip().startswith("/*") and len(parts[0].strip()) in (6, 7):
                t = parts[1].strip().split()
                if t:
                    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
                    ops[op.split(".")[0].rstrip(";")] += 1
        extra = " ".join(f"{o}:{ops[o]}" for o in ("UBLKCP", "UGETNEXTWORKID", "SYNCS", "LDGSTS") if ops[o])
        f.write(f"{sum(ops.values()):6d}  {k}\n        " + " ".join(f"{o}:{c}" for o, c in ops.most_common(14)) +
                (f"   | sm_100 / async: {extra}" if extra else "") + "\n")
for k, lines in blocks.items():
    if "lmfb_k1ILi1ELb1ELi5ELi3ELb0ELb0" in k or "lmfb_k1ILi1ELb0ELi5ELi3ELb0ELb0" in k:      # the default reim kernels
        short = "k1_bwd_reim_w5" if "Lb1ELi5" in k else "k1_fwd_reim_w5"
        with open(os.path.join(P, f"{tag}_sass_{short}.txt"), "w") as f:
            f.write(f"# {k}\n")
            import re as _re
            keep = []
            for l in lines:
                if _re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l):          # instruction line: drop the hex encoding
                    keep.append(_re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip())
                elif l.strip().startswith(".") or "headerflags" in l:
                    keep.append(l.rstrip())
            f.write("\n".join(keep) + "\n")
print("saved", sorted(x for x in os.listdir(P) if x.startswith(tag) or x == "traffic.json"))
