#!/bin/bash
# quick GPU iteration: parity tests, two bench lines, one ncu --set full capture of K1 on the sweep workload
TAG=${1:-q}
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_$TAG.txt
echo "== bench chime"; timeout 600 python bench.py --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; python - <<PY
import json
for f in ("gpurun_out/bench_$TAG.json",):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("value %.3e audio-s/s  ms/step %.4f  dom frac %.3f k1f frac %.3f step frac %.3f e2e %.3e" % (d["value"], d["ms_per_step"], r["frac"], r["k1_fwd_frac"], r["step_frac"], d["e2e"]["value"]), r["kernels_ms"])
    except Exception as e: print("bench parse failed", e)
PY
echo "== bench sweep"; timeout 600 python bench.py --workload sweep_256x10s --steps 40 --warmup 5 --no-cpu > gpurun_out/bench_sweep_$TAG.json 2> gpurun_out/bench_sweep_$TAG.err; tail -3 gpurun_out/bench_sweep_$TAG.err; python - <<PY
import json
for f in ("gpurun_out/bench_sweep_$TAG.json",):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("value %.3e audio-s/s  ms/step %.4f  dom frac %.3f k1f frac %.3f step frac %.3f e2e %.3e" % (d["value"], d["ms_per_step"], r["frac"], r["k1_fwd_frac"], r["step_frac"], d["e2e"]["value"]), r["kernels_ms"])
    except Exception as e: print("bench parse failed", e)
PY
if [ "$2" != "noncu" ]; then
echo "== ncu full (sweep)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lmfb_k1 -s 6 -c 2 -o gpurun_out/prof_sweep_$TAG -f python bench.py --workload sweep_256x10s --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/prof_sweep_$TAG.ncu-rep
fi
