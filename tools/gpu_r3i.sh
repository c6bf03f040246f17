#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_libs.sh "base pfn nopf base" sweep_256x10s
bash tools/gpu_libs.sh "base pfn nopf" chime4_30x6s
