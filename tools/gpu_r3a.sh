#!/bin/bash
# round-2 session, run A: new parity tests + the new bench line (roofline_large, paired)
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_r3a.txt
echo "== bench default"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r3a.json 2> gpurun_out/bench_r3a.err; tail -3 gpurun_out/bench_r3a.err; cat gpurun_out/bench_r3a.json
echo "== bench paired"; timeout 600 python bench.py --workload paired_30x6s --steps 20 --warmup 5 --no-large > gpurun_out/bench_paired_r3a.json 2> gpurun_out/bench_paired_r3a.err; tail -3 gpurun_out/bench_paired_r3a.err; cat gpurun_out/bench_paired_r3a.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r3a.json 2>&1; cat gpurun_out/bench_ref_r3a.json
