#!/bin/bash
# On the GPU box: fused pass vs layers (parity + time), then the cycle accounting of the instrumented build.  $1 = log tag
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 200 python tools/mega_check.py ${2:-64} 5 > gpurun_out/mega_check_$1.log 2>&1
FSUAE_LIB_PATH=$PWD/fs_uae_image_enhancer_project_b200/libfsuae_timing.so timeout 120 python tools/mega_timing.py ${2:-64} > gpurun_out/mega_timing_$1.log 2>&1
cat gpurun_out/mega_check_$1.log gpurun_out/mega_timing_$1.log
