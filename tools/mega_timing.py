#!/usr/bin/env python3
"""Debug aid: where the engines of the single fused pass (csrc/mega.cu) spend their cycles.

Needs a library built with FSUAE_EXTRA_NVCC_FLAGS=-DFSUAE_EPI_TIMING (point FSUAE_LIB_PATH at it).  Counters come from the
probe CTA of each stage (team 0, middle strip, leader of the pair); see the table in mega.cu.

    python tools/mega_timing.py [frames] [u8|f32]
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
fmt = sys.argv[2] if len(sys.argv) > 2 else "u8"
if fmt == "u8":
    x = (torch.randint(0, 16, (b, 576, 752, 4), dtype=torch.uint8) * 17).to(dev)
    run = m.forward_framebuffer
else:
    x = torch.rand(b, 3, 576, 752, device=dev)
    run = m.forward
run(x)
lib = _lib.load()
out = (C.c_ulonglong * (8 * 16))()
lib.fsuae_debug_mega_timing(out, 1)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
run(x)
ev1.record()
torch.cuda.synchronize()
lib.fsuae_debug_mega_timing(out, 1)
print(f"{b} frames: {ev0.elapsed_time(ev1) * 1e3 / b:.1f} us/frame (variant {m.engine_for(dev, 576, 752).variant})")
print("cycles per row of the probe CTA (team 0, middle strip, leader)")
print("layer | issuer: total  in-rows  tempty  issue | producer: total  smem-slot  upstream | epilogue warp: total  backpr  tfull  work [fence release] (rows)")
for layer in range(7):
    v = [out[layer * 16 + i] for i in range(16)]
    ir, pr, er = max(v[13], 1), max(v[12], 1), max(v[11], 1)
    print(f"conv{layer + 1} | {v[0] / ir:8.0f} {v[1] / ir:8.0f} {v[2] / ir:8.0f} {v[3] / ir:7.0f} | {v[6] / pr:8.0f} {v[4] / pr:8.0f} {v[5] / pr:8.0f} |"
          f" {v[10] / er:8.0f} {v[7] / er:8.0f} {v[8] / er:8.0f} {v[9] / er:8.0f} [{v[14] / er:6.0f} {v[15] / er:6.0f}] ({v[11]})")
h = [out[7 * 16 + i] for i in range(16)]
hr = max(h[4], 1)
print("producer credit step (wait for the copy of fill - 2), cycles per row: " + "  ".join(f"conv{l + 1} {h[5 + l] / max(out[l * 16 + 12], 1):.0f}" for l in range(7)))
print(f"conv4 issuer: waits for the peer's relay {h[12] / max(out[3 * 16 + 13], 1):.0f} cycles per row")
print(f"head worker 0 (stage B): {h[4]} rows, {h[13] / hr:.0f} cycles per row in all")
print(f"head worker 0 (stage B) per row: back-pressure {h[1] / hr:.0f}  wait + LUT + stores {h[2] / hr:.0f}  barrier + release {h[3] / hr:.0f}  prefetch {h[0] / hr:.0f}")
