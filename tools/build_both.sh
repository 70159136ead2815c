#!/bin/bash
# Build the production library and the instrumented one (-DFSUAE_EPI_TIMING -> libfsuae_timing.so) side by side.
cd "$(dirname "$0")/.." || exit 1
(FSUAE_LIB_PATH=$PWD/fs_uae_image_enhancer_project_b200/libfsuae_timing.so FSUAE_EXTRA_NVCC_FLAGS="-DFSUAE_EPI_TIMING $FSUAE_TIMING_EXTRA" python -m fs_uae_image_enhancer_project_b200.build 2>&1 | grep -v Warning | tail -5) &
python -m fs_uae_image_enhancer_project_b200.build 2>&1 | grep -v Warning | tail -5
wait
