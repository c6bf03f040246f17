#!/bin/bash
# N-GPU bench (torchrun, one rank per GPU) + the reference arm.  usage: tools/gpu_multi.sh N tag
N=${1:-2}; TAG=${2:-m}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 400 --warmup 10 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err
tail -2 gpurun_out/bench_${N}gpu_$TAG.err; cat gpurun_out/bench_${N}gpu_$TAG.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload sweep_256x10s --steps 40 --warmup 5 > gpurun_out/bench_${N}gpu_sweep_$TAG.json 2> gpurun_out/bench_${N}gpu_sweep_$TAG.err
tail -2 gpurun_out/bench_${N}gpu_sweep_$TAG.err; cat gpurun_out/bench_${N}gpu_sweep_$TAG.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 10 --warmup 2 > gpurun_out/bench_ref_${N}gpu_$TAG.json 2> gpurun_out/bench_ref_${N}gpu_$TAG.err
cat gpurun_out/bench_ref_${N}gpu_$TAG.json
