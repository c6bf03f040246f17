#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_r3r.txt
echo "== grad wave"; timeout 600 python tools/time_grad_wave.py 2>&1 | grep "warps_bwd=0"
