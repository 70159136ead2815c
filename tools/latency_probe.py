#!/usr/bin/env python3
"""Single-frame latency (p50 over 300 enqueues, CUDA events) of the bf16 flagship; honours the FSUAE_* debugging switches."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs_uae_image_enhancer_project_b200 import model_pix_shuffle, _lib
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
m = model_pix_shuffle.get_model("lightweight").to(dev).set_precision("bf16")
m.chunk_frames = max(n, 1)
x = torch.randint(0, 256, (n, 576, 752, 4), dtype=torch.uint8, device=dev)
out = torch.empty_like(x)
eng = m.engine_for(dev, 576, 752)
flags = _lib.FLAG_GAMMA_IN | _lib.FLAG_GAMMA_OUT
for _ in range(30):
    eng.enqueue(x, out, n, _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags)
torch.cuda.synchronize()
lat = []
for _ in range(300):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.enqueue(x, out, n, _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags); b.record(); b.synchronize()
    lat.append(a.elapsed_time(b) * 1e3)
lat.sort()
print(f"frames {n}: p50 {lat[150]:.1f} us  p10 {lat[30]:.1f}  p99 {lat[296]:.1f}")
