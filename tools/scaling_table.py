#!/usr/bin/env python3
"""Copy the multi-GPU bench lines of one round from gpurun_out/scaling/ into profiles/scaling/ and write
the scaling tables (profiles/scaling/README.md).   usage: tools/scaling_table.py <tag>"""
import glob
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "scaling")
DST = os.path.join(ROOT, "profiles", "scaling")
tag = sys.argv[1]
os.makedirs(DST, exist_ok=True)
lines = {}
for path in sorted(glob.glob(os.path.join(SRC, f"{tag}_*"))):
    name = os.path.basename(path)
    if name.endswith(".err"):
        continue
    shutil.copy(path, os.path.join(DST, name))
    if name.endswith(".json"):
        try:
            lines[name[len(tag) + 1:-5]] = json.loads(open(path).read().strip().splitlines()[-1])
        except Exception:
            pass


def row(prefix, key=lambda d: d["value"], fmt="%.3e"):
    cells, base = [], None
    for n in (1, 2, 4, 8):
        d = lines.get(f"{prefix}_n{n}")
        if d is None:
            cells.append("—")
            continue
        v = key(d)
        if n == 1:
            base = v
        eff = (" (%.2f)" % (v / (n * base))) if base and n > 1 else ""
        cells.append((fmt % v) + eff)
    return cells


out = [f"# Multi-GPU measurements, tag {tag} (one box, 8 x B200, one process per GPU, no data-path collective)",
       "",
       "Weak scaling: every rank owns its own utterances; `value` = audio-seconds of all ranks / the slowest rank's",
       "device time (CUDA events, max over ranks).  In brackets: efficiency against N x the 1-GPU value.",
       "e2e = the same through the public API with pinned HOST buffers (PCIe inside the timed region).",
       "",
       "| workload (per GPU) | quantity | 1 GPU | 2 | 4 | 8 |", "|---|---|---|---|---|---|"]
for prefix, label in (("chime", "config #2: 30 x 6 s, fwd+bwd"), ("paired", "configs[3]: paired 30 x 6 s (noisy fwd+bwd + clean fwd)"),
                      ("sweep256", "256 x 10 s, fwd+bwd"), ("aas", "configs[2]: full AAS step, 30 x 6 s")):
    if not any(k.startswith(prefix + "_") for k in lines):
        continue
    out.append("| %s | audio-s/s | %s |" % (label, " | ".join(row(prefix))))
    out.append("| | ms / step | %s |" % " | ".join(row(prefix, lambda d: d["ms_per_step"], "%.4f")).replace(" (", " <!-- (").replace(")", ") -->"))
    if prefix != "sweep256":
        out.append("| | e2e audio-s/s | %s |" % " | ".join(row(prefix, lambda d: d["e2e"]["value"])))
aas = {n: lines.get(f"aas_n{n}") for n in (1, 2, 4, 8)}
if any(aas.values()):
    out += ["", "## The AAS step (bench_aas.py): where the time goes", "",
            "| GPUs | ms / step | ms without the all-reduce | all-reduce in the step (ms, incl. waiting for the slowest rank) | all-reduce by itself (ms) | bus GB/s | front-end by itself (ms) | front-end share |",
            "|---|---|---|---|---|---|---|---|"]
    for n, d in aas.items():
        if d is None:
            continue
        a = d["aas_step"]
        out.append("| %d | %.1f | %.1f | %.2f | %s | %s | %.3f | %.2f %% |" % (
            n, a["ms_step"], a["ms_step_without_allreduce"], a["ms_allreduce_in_step"],
            "%.2f" % a["ms_allreduce_isolated"] if a.get("ms_allreduce_isolated") else "—",
            "%.0f" % a["allreduce_busbw_gbs"] if a.get("allreduce_busbw_gbs") else "—",
            a["ms_frontend_isolated"], 100 * a["frontend_share_of_step"]))
    d = next(v for v in aas.values() if v)
    out += ["", "Gradient buffer: %.0f MB fp32 (G %d + D %d + ASR %d parameters), one NCCL all-reduce per step; "
            "SyncBatchNorm in the acoustic model when sharded." % (d["aas_step"]["allreduce_bytes"] / 1e6, d["aas_step"]["params"]["G"],
                                                                  d["aas_step"]["params"]["D"], d["aas_step"]["params"]["ASR"])]
for n in (1, 8):
    p = os.path.join(SRC, f"{tag}_pcie_n{n}.json")
    if os.path.exists(p):
        try:
            d = json.loads(open(p).read().strip().splitlines()[-1])
        except Exception:
            continue
        if n == 1:
            out += ["", "## Host-fabric ceiling of the e2e numbers (tools/pcie_ceiling.py: the same bytes, no kernels)", "",
                    "| GPUs | H2D alone GB/s per rank | D2H alone | both at once: H2D | D2H | total GB/s | e2e ceiling, config #2 (audio-s/s) |",
                    "|---|---|---|---|---|---|---|"]
        c = d["copies"]
        out.append("| %d | %.1f | %.1f | %.1f | %.1f | %.0f | %.3e |" % (
            n, c["h2d_only"]["h2d_gbs_per_rank"], c["d2h_only"]["d2h_gbs_per_rank"], c["both"]["h2d_gbs_per_rank"],
            c["both"]["d2h_gbs_per_rank"], d["total_gbs_both_directions"], d["e2e_ceiling_audio_s_per_s"]))
sw = os.path.join(SRC, f"{tag}_sweep_n8.md")
if os.path.exists(sw):
    out += ["", "## configs[4] sweep at 8 GPUs (tools/sweep.py --quick; the full 1-GPU grid is profiles/%s_sweep_batch_x_seconds.md)" % tag, ""]
    out += open(sw).read().splitlines()
open(os.path.join(DST, "README.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
