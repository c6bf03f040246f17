#!/bin/bash
# A/B several prebuilt libraries (aas_enhancement_b200/libaas_lmfb_<tag>.so) on the sweep workload.
# usage: tools/gpu_libs.sh "<tag list>" [workload]
WL=${2:-sweep_256x10s}
for t in $1; do
  timeout 300 python bench.py --lib aas_enhancement_b200/libaas_lmfb_$t.so --workload $WL --steps 100 --warmup 5 --no-cpu --no-e2e --no-large > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/v.json")); r=d["roofline"]; k=r["kernels_ms"]
    print("lib=%-10s %-14s value %.3e ms/step %.4f step_frac %.3f  k1f %.4f ms (frac %.3f)  k1b %.4f ms (frac %.3f) k2f %.4f k2b %.4f" % ("$t", "$WL", d["value"], d["ms_per_step"], r["step_frac"], k["k1_fwd"], r["k1_fwd_frac"], k["k1_bwd"], r["k1_bwd_frac"], k["k2_fwd"], k["k2_bwd"]))
except Exception as e: print("failed", e)
PY
done
