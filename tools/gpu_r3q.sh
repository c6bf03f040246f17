#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (wave gradient tests)"; timeout 900 python -m pytest tests -m gpu -x -q -k "wave_grad or strided_wave" 2>&1 | tail -4
echo "== grad wave"; timeout 600 python tools/time_grad_wave.py 2>&1 | tail -18
