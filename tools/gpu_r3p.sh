#!/bin/bash
mkdir -p gpurun_out
bash tools/sanitize.sh 2>&1 | tail -60
bash tools/sanitize_big.sh 2>&1 | tail -30
