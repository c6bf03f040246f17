#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_libs.sh "base pfs k2r base" sweep_256x10s
bash tools/gpu_libs.sh "pfs k2r" chime4_30x6s
