#!/bin/bash
# run B: gather phase 3 + generic/multi-channel paths: tests, bench (chime + large), timelines
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_r3b.txt
echo "== bench default"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r3b.json 2> gpurun_out/bench_r3b.err; tail -3 gpurun_out/bench_r3b.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r3b.json"))
for key in ("roofline","roofline_large"):
    r=d[key]; print(key, "ms/step %.4f step_frac %.3f k1f_frac %.3f k1b_frac %.3f" % (r["ms_per_step"], r["step_frac"], r["k1_fwd_frac"], r["k1_bwd_frac"]), r["kernels_ms"])
print("value %.3e e2e %.3e launches/step %s" % (d["value"], d["e2e"]["value"], d["gpu_launches_per_step"]))
PY
echo "== timeline chime W=4"; TL_W=4 timeout 300 python tools/timeline.py chime 2>&1 | tail -22
echo "== timeline sweep W=3"; TL_W=3 timeout 300 python tools/timeline.py sweep 2>&1 | tail -22
echo "== reference arm x3"
for i in 1 2 3; do timeout 300 python bench.py --impl reference --steps 20 --warmup 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['cpu_baseline']['sample'])"; done
