#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
for key in ("roofline","roofline_large"):
    if key in d:
        r=d[key]; print(key, "ms/step %.4f step_frac %.3f k1f_frac %.3f k1b_frac %.3f" % (r["ms_per_step"], r["step_frac"], r["k1_fwd_frac"], r["k1_bwd_frac"]), {k: round(v,4) for k,v in r["kernels_ms"].items()}, d["gpu_launches_per_step"])
PY
}
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_r3l.txt
echo "== bench default (pdl)"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench_r3l.json 2> gpurun_out/bench_r3l.err; tail -3 gpurun_out/bench_r3l.err; show gpurun_out/bench_r3l.json
echo "== bench default (no pdl)"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-pdl > gpurun_out/bench_r3l_np.json 2> gpurun_out/bench_r3l.err; tail -3 gpurun_out/bench_r3l.err; show gpurun_out/bench_r3l_np.json
echo "== bench paired (pdl)"; timeout 600 python bench.py --workload paired_30x6s --steps 20 --warmup 5 --no-cpu --no-e2e --no-large > gpurun_out/bench_r3l_p.json 2> gpurun_out/bench_r3l.err; show gpurun_out/bench_r3l_p.json
