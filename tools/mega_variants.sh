#!/bin/bash
# On the GPU box: time every libfsuae_var*.so experiment build against the layer kernels.  $1 = log tag
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
for so in fs_uae_image_enhancer_project_b200/libfsuae_var*.so; do
  echo "== $so"
  FSUAE_LIB_PATH=$PWD/$so timeout 200 python tools/mega_check.py ${2:-64} ${3:-20} 2>&1 | grep -v "^layers"
done 2>&1 | tee gpurun_out/mega_var_$1.log
