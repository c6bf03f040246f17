#!/bin/bash
# Run on the B200 box via gpurun: GPU parity tests, smoke, a bench line, and ncu captures.
# usage: tools/gpu_check.sh [tag]
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi_$TAG.txt 2>&1
nproc > gpurun_out/host_$TAG.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> gpurun_out/host_$TAG.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_$TAG.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke_$TAG.txt
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
echo "== bench sweep"; timeout 600 python bench.py --workload sweep_256x10s --steps 40 --warmup 5 --no-cpu > gpurun_out/bench_sweep_$TAG.json 2> gpurun_out/bench_sweep_$TAG.err; tail -3 gpurun_out/bench_sweep_$TAG.err; cat gpurun_out/bench_sweep_$TAG.json
echo "== bench paired"; timeout 600 python bench.py --workload paired_30x6s --no-large > gpurun_out/bench_paired_$TAG.json 2> gpurun_out/bench_paired_$TAG.err; tail -3 gpurun_out/bench_paired_$TAG.err; cat gpurun_out/bench_paired_$TAG.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$TAG.json 2>&1; cat gpurun_out/bench_ref_$TAG.json
echo "== grad wave / modes"; (timeout 300 python tools/time_grad_wave.py; timeout 300 python tools/time_modes.py) 2>&1 | tee gpurun_out/modes_$TAG.txt
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-large > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lmfb_k1 -s 6 -c 2 -o gpurun_out/prof_chime_$TAG -f python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-large > gpurun_out/ncu_full_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lmfb_k1 -s 6 -c 2 -o gpurun_out/prof_sweep_$TAG -f python bench.py --workload sweep_256x10s --steps 4 --warmup 3 --no-cpu --no-e2e --no-large >> gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -20
