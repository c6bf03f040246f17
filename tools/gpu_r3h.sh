#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
for key in ("roofline","roofline_large"):
    if key in d:
        r=d[key]; print(key, "ms/step %.4f step_frac %.3f k1f_frac %.3f k1b_frac %.3f" % (r["ms_per_step"], r["step_frac"], r["k1_fwd_frac"], r["k1_bwd_frac"]), {k: round(v,4) for k,v in r["kernels_ms"].items()})
PY
}
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_r3h.txt
echo "== bench default"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench_r3h.json 2> gpurun_out/bench_r3h.err; tail -3 gpurun_out/bench_r3h.err; show gpurun_out/bench_r3h.json
echo "== instruction counts"
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:lmfb_k1 -s 6 -c 2 python bench.py --workload sweep_256x10s --steps 4 --warmup 3 --no-cpu --no-e2e --no-large 2>&1 | grep -E "lmfb_k1|inst_executed|duration" | head -8
