#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_r3c.txt
echo "== bench default"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r3c.json 2> gpurun_out/bench_r3c.err; tail -3 gpurun_out/bench_r3c.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r3c.json"))
for key in ("roofline","roofline_large"):
    r=d[key]; print(key, "ms/step %.4f step_frac %.3f k1f_frac %.3f k1b_frac %.3f" % (r["ms_per_step"], r["step_frac"], r["k1_fwd_frac"], r["k1_bwd_frac"]), r["kernels_ms"])
print("value %.3e e2e %.3e launches/step %s" % (d["value"], d["e2e"]["value"], d["gpu_launches_per_step"]))
PY
