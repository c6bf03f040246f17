#!/bin/bash
mkdir -p gpurun_out
for w in "8 8" "5 8" "8 5" "6 6"; do set -- $w
timeout 300 python bench.py --warps-fwd $1 --warps-bwd $2 --steps 20 --warmup 5 --no-cpu --no-e2e --no-large > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open("gpurun_out/v.json")); r=d["roofline"]; k=r["kernels_ms"]
print("wf=%s wb=%s chime ms/step %.4f value %.3e k1f %.4f k1b %.4f k2f %.4f k2b %.4f" % (sys.argv[1], sys.argv[2], d["ms_per_step"], d["value"], k["k1_fwd"], k["k1_bwd"], k["k2_fwd"], k["k2_bwd"]))
PY
done
