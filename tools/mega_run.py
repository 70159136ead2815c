#!/usr/bin/env python3
"""Debug aid: one fused-pass launch on synthetic framebuffers (for ncu captures): python tools/mega_run.py [frames] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs_uae_image_enhancer_project_b200 import model_pix_shuffle  # noqa: E402

dev = torch.device("cuda", 0)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(3)
m = model_pix_shuffle.get_model("lightweight").to(dev).set_precision("bf16")
m.chunk_frames = b
fb = (torch.randint(0, 16, (b, 576, 752, 4), dtype=torch.uint8) * 17).to(dev)
for _ in range(reps):
    y = m.forward_framebuffer(fb)
torch.cuda.synchronize()
print("ok", m.engine_for(dev, 576, 752).variant, int(y.sum()))
