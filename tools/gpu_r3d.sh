#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_r3d.txt
for wf in 5 4 6; do
echo "== bench default W=$wf"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --warps-fwd $wf --warps-bwd $wf > gpurun_out/bench_r3d_$wf.json 2> gpurun_out/bench_r3d.err; tail -3 gpurun_out/bench_r3d.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r3d_$wf.json"))
for key in ("roofline","roofline_large"):
    r=d[key]; print(key, "ms/step %.4f step_frac %.3f k1f_frac %.3f k1b_frac %.3f" % (r["ms_per_step"], r["step_frac"], r["k1_fwd_frac"], r["k1_bwd_frac"]), r["kernels_ms"])
PY
done
echo "== timeline chime W=5"; TL_W=5 timeout 300 python tools/timeline.py chime 2>&1 | tail -26
echo "== timeline sweep W=5"; TL_W=5 timeout 300 python tools/timeline.py sweep 2>&1 | tail -26
