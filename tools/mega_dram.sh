#!/bin/bash
# On the GPU box: DRAM bytes + duration of one 64-frame fused pass (production build; FSUAE_* switches pass through), per library
cd "$(dirname "$0")/.." || exit 1
for so in fs_uae_image_enhancer_project_b200/libfsuae_enhancer.so fs_uae_image_enhancer_project_b200/libfsuae_var*.so; do
  [ -f "$so" ] || continue
  echo "== $so $FSUAE_NO_L2_WINDOW"
  FSUAE_LIB_PATH=$PWD/$so timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fused_pass -s 2 -c 1 python tools/mega_run.py 64 4 2>&1 | grep -E "dram__|gpu__time"
done 2>&1 | tee -a gpurun_out/mega_dram_$1.log
