#!/bin/bash
# compare warps-per-tile variants on the sweep workload (and chime) -- no ncu
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for wf in 3 4 2 5; do for wb in 4 3 2; do
  if [ "$wf" != "3" ] && [ "$wb" != "4" ]; then continue; fi
  for wl in sweep_256x10s chime4_30x6s; do
    timeout 300 python bench.py --warps-fwd $wf --warps-bwd $wb --workload $wl --steps 200 --warmup 5 --no-cpu --no-e2e --no-large > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/v.json")); r=d["roofline"]; k=r["kernels_ms"]
    print("wf=$wf wb=$wb %-14s value %.3e step_frac %.3f  k1f %.4f ms (frac %.3f)  k1b %.4f ms (frac %.3f)" % ("$wl", d["value"], r["step_frac"], k["k1_fwd"], r["k1_fwd_frac"], k["k1_bwd"], r["frac"]))
except Exception as e: print("failed", e)
PY
  done
done; done
