#!/usr/bin/env python3
"""Debug aid: cycles every CTA of the single fused pass took (entry -> its last engine's last row), by stage.

Needs a library built with FSUAE_EXTRA_NVCC_FLAGS=-DMG_DBG_FINISH (point FSUAE_LIB_PATH at it); with -DMG_DBG_CUT=0x7F on top
the engines run free of each other and the figures are each stage's stand-alone run time.

    python tools/mega_finish.py [frames]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs_uae_image_enhancer_project_b200 import _lib, model_pix_shuffle  # noqa: E402

dev = torch.device("cuda", 0)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = model_pix_shuffle.get_model("lightweight").to(dev).set_precision("bf16")
m.chunk_frames = b
x = (torch.randint(0, 16, (b, 576, 752, 4), dtype=torch.uint8) * 17).to(dev)
for _ in range(3):
    m.forward_framebuffer(x)
lib = _lib.load()
out = (C.c_longlong * 160)()
lib.fsuae_debug_mega_finish(out)
rows = 288 * ((b + 1) // 2) / 6
names = ["A conv5+conv2", "B conv4+head", "C conv3+conv7", "D conv6+conv1"]
for st in range(4):
    v = [out[blk] for blk in range(144) if (blk >> 1) & 3 == st]
    print(f"stage {names[st]}: cycles per row min {min(v) / rows:7.0f}  mean {sum(v) / len(v) / rows:7.0f}  max {max(v) / rows:7.0f}")

# entry / exit of every CTA on the global timer, last four back-to-back launches: how far apart the CTAs of a launch start,
# how long the launch lives, and the gap to the next one
for _ in range(8):
    m.forward_framebuffer(x)
gt = (C.c_ulonglong * (4 * 2 * 160))()
nl = C.c_uint(0)
lib.fsuae_debug_mega_globaltimer(gt, C.byref(nl))
n = nl.value
prev_end = None
for k in range(n - 4, n):
    sl = k & 3
    st = [gt[(sl * 2 + 0) * 160 + i] for i in range(144)]
    en = [gt[(sl * 2 + 1) * 160 + i] for i in range(144)]
    gap = f"  gap after the previous launch {(min(st) - prev_end) / 1e3:7.1f} us" if prev_end else ""
    print(f"launch {k}: CTA entries spread over {(max(st) - min(st)) / 1e3:6.1f} us, first entry -> last exit {(max(en) - min(st)) / 1e3:8.1f} us, "
          f"exits spread over {(max(en) - min(en)) / 1e3:6.1f} us{gap}")
    prev_end = max(en)

# the probe CTAs (team 0, middle strip, leader) of the last launch: when each engine finished its first and its last row,
# relative to the launch's first CTA entry -- the fill and the drain of the pipeline, hop by hop
rows_t = (C.c_ulonglong * 16)()
lib.fsuae_debug_mega_rows(rows_t)
sl = (n - 1) & 3
t0 = min(gt[(sl * 2 + 0) * 160 + i] for i in range(144))
t1 = max(gt[(sl * 2 + 1) * 160 + i] for i in range(144))
names = ["head"] + [f"conv{i}" for i in range(1, 8)]
print("engine : first row done at / last row done before the end of the launch (us)")
for i, nm in enumerate(names):
    print(f"{nm:6s} : {(rows_t[i] - t0) / 1e3:8.1f}   {(t1 - rows_t[8 + i]) / 1e3:8.1f}")
