#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_libs.sh "base wbr base wbr" sweep_256x10s
bash tools/gpu_libs.sh "base wbr" chime4_30x6s
