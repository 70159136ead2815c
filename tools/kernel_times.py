#!/usr/bin/env python3
"""Per-kernel device times of one forward (engine profiling mode): python tools/kernel_times.py FAMILY PRESET [PREC] [BATCH]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs_uae_image_enhancer_project_b200 import model_pix_shuffle, model_conv3, model_conv5
fam, preset = sys.argv[1], sys.argv[2]
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
b = int(sys.argv[4]) if len(sys.argv) > 4 else 16
dev = torch.device("cuda", 0)
mod = {"pix_shuffle": model_pix_shuffle, "conv3": model_conv3, "conv5": model_conv5}[fam]
m = mod.get_model(preset).to(dev).set_precision(prec)
m.chunk_frames = b
x = torch.randint(0, 256, (b, 4, 576, 752), dtype=torch.uint8, device=dev) if fam == "conv3" else torch.rand(b, 3, 576, 752, device=dev)
for _ in range(2):
    m(x)
eng = m.engine_for(dev, 576, 752)
eng.set_profiling(True)
m(x)
torch.cuda.synchronize()
tot = 0
for label, ms in eng.kernel_times():
    print(f"{label:28s} {1e3 * ms / b:9.1f} us/frame")
    tot += ms
print(f"{'total':28s} {1e3 * tot / b:9.1f} us/frame")
