#!/usr/bin/env python3
"""Soak test of the single fused pass: many back-to-back passes under concurrent host <-> device copies, every result compared
bit by bit with the first (the rings are re-used every few microseconds and guarded only by flags).

    python tools/mega_soak.py [passes] [frames]
"""
import os
import sys
import threading

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs_uae_image_enhancer_project_b200 import model_pix_shuffle  # noqa: E402

dev = torch.device("cuda", 0)
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.manual_seed(11)
m = model_pix_shuffle.get_model("lightweight").to(dev).set_precision("bf16")
m.chunk_frames = b
g = torch.Generator().manual_seed(2)
x = (torch.randint(0, 16, (b, 576, 752, 4), generator=g, dtype=torch.uint8) * 17).to(dev)
h = torch.empty_like(x, device="cpu").pin_memory()
d = torch.empty_like(x)
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
stop = False


def copier():
    while not stop:
        with torch.cuda.stream(s_up):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s_dn):
            h.copy_(d, non_blocking=True)
        s_up.synchronize()
        s_dn.synchronize()


first = m.forward_framebuffer(x).clone()
assert m.engine_for(dev, 576, 752).last_launch_count == 1
th = threading.Thread(target=copier)
th.start()
bad = 0
try:
    for i in range(passes):
        y = m.forward_framebuffer(x)
        if i % 8 == 7 and not torch.equal(y, first):
            bad += 1
finally:
    stop = True
    th.join()
torch.cuda.synchronize()
print(f"{passes} passes of {b} frames under copy traffic: {bad} of {passes // 8} compared results differ")
sys.exit(1 if bad else 0)
