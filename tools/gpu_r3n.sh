#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_r3n.txt
echo "== grad wave"; timeout 300 python tools/time_grad_wave.py 2>&1 | tail -5
echo "== modes"; timeout 300 python tools/time_modes.py 2>&1 | tail -8
echo "== bench default (full line)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r3n.json 2> gpurun_out/bench_r3n.err; tail -3 gpurun_out/bench_r3n.err; cat gpurun_out/bench_r3n.json
