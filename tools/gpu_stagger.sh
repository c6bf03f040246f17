#!/bin/bash
for st in 0 1500 3000 5000 8000; do
  for wl in chime4_30x6s sweep_256x10s; do
    AAS_LMFB_STAGGER_NS=$st timeout 300 python bench.py --workload $wl --steps 200 --warmup 5 --no-cpu > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/v.json")); r=d["roofline"]; k=r["kernels_ms"]
    print("stagger=$st %-14s value %.3e step_frac %.3f  k1f %.4f ms  k1b %.4f ms" % ("$wl", d["value"], r["step_frac"], k["k1_fwd"], k["k1_bwd"]))
except Exception as e: print("failed", e)
PY
  done
done
