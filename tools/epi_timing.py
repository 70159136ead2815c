#!/usr/bin/env python3
"""Debug aid: cycle breakdown of the run-time epilogue (library built with FSUAE_EXTRA_NVCC_FLAGS=-DFSUAE_EPI_TIMING).
Counters of block 0 / warp 2 / lane 0: [0] tfull wait, [1] rows, [2] parameter loads, [3] tcgen05.ld + wait,
[4] bias add, [5] activation + pack + store, [6] whole row, [7] chunks."""
import os, sys, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs_uae_image_enhancer_project_b200 import model_conv3, model_conv5, model_pix_shuffle, _lib
fam, preset = sys.argv[1], sys.argv[2]
b = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dev = torch.device("cuda", 0)
mod = {"pix_shuffle": model_pix_shuffle, "conv3": model_conv3, "conv5": model_conv5}[fam]
m = mod.get_model(preset).to(dev).set_precision("bf16")
m.chunk_frames = b
x = torch.randint(0, 256, (b, 4, 576, 752), dtype=torch.uint8, device=dev) if fam == "conv3" else torch.rand(b, 3, 576, 752, device=dev)
m(x)
lib = _lib.load()
out = (C.c_ulonglong * 8)()
lib.fsuae_debug_epi_timing(out, 1)
m(x)
lib.fsuae_debug_epi_timing(out, 1)
v = list(out)
rows, chunks = max(v[1], 1), max(v[7], 1)
print(f"rows {v[1]} chunks {v[7]}")
print(f"per row : tfull wait {v[0]/rows:8.0f}  whole {v[6]/rows:8.0f} cycles")
print(f"per chunk: params {v[2]/chunks:7.0f}  tmem ld {v[3]/chunks:7.0f}  bias {v[4]/chunks:7.0f}  act+store {v[5]/chunks:7.0f}")
