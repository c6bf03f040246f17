#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_r3u.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench (driver-style)"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r3u.json 2> gpurun_out/bench_r3u.err; tail -2 gpurun_out/bench_r3u.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r3u.json"))
print("value %.4e ms/step %.4f e2e %.4e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
for k in ("roofline","roofline_large"):
    r=d[k]; print(k, "step_frac %.3f frac %.3f" % (r["step_frac"], r["frac"]), {a: round(b,4) for a,b in r["kernels_ms"].items()})
PY
echo "== full sweep, 1 GPU"; timeout 900 python tools/sweep.py --cpu-only 2>&1 | tail -8; timeout 1200 python tools/sweep.py gpurun_out/r3_sweep_batch_x_seconds.md 2>&1 | tail -75
