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
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_r3g.txt
echo "== bench default"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/bench_r3g.json 2> gpurun_out/bench_r3g.err; tail -3 gpurun_out/bench_r3g.err; show gpurun_out/bench_r3g.json
echo "== timeline sweep W=5"; TL_W=5 timeout 300 python tools/timeline.py sweep 2>&1 | grep -E "CTAs|warp 0|detail" | head -8
echo "== ncu full (sweep)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lmfb_k1 -s 6 -c 2 -o gpurun_out/prof_sweep_r3g -f python bench.py --workload sweep_256x10s --steps 4 --warmup 3 --no-cpu --no-e2e --no-large > gpurun_out/ncu_full_r3g.log 2>&1
ls -la gpurun_out/prof_sweep_r3g.ncu-rep
