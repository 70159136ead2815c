#!/bin/bash
# On the GPU box: cycle accounting (tools/mega_timing.py) of every instrumented experiment build libfsuae_t*.so.  $1 = log tag
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
for so in fs_uae_image_enhancer_project_b200/libfsuae_t*.so; do
  echo "== $so"
  FSUAE_LIB_PATH=$PWD/$so timeout 120 python tools/mega_timing.py ${2:-64} 2>&1 | tail -14
done 2>&1 | tee gpurun_out/mega_tvar_$1.log
