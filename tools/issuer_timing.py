#!/usr/bin/env python3
"""Debug aid (library built with FSUAE_EXTRA_NVCC_FLAGS=-DFSUAE_EPI_TIMING): where MMA issuer 0 of CTA 0 spends its cycles,
summed over the resident-weight layer kernels of one flagship pass.  Per input row it handles:
[0] ring-row waits  [1] accumulator (tempty) wait  [2] token wait  [3] MMA issue block  [4] commits + hand-over  [7] rows."""
import os, sys, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs_uae_image_enhancer_project_b200 import model_pix_shuffle, _lib
dev = torch.device("cuda", 0)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = model_pix_shuffle.get_model("lightweight").to(dev).set_precision("bf16")
m.chunk_frames = b
x = torch.rand(b, 3, 576, 752, device=dev)
m(x)
lib = _lib.load()
out = (C.c_ulonglong * 32)()
lib.fsuae_debug_epi_timing(out, 1)
m(x)
lib.fsuae_debug_epi_timing(out, 1)
full = list(out); v = full[16:24] + full[8:16]; rows = max(v[7], 1)
print(f"rows handled by issuer 0 of CTA 0: {v[7]}")
print(f"         of the ring waits, own CTA's TMA barrier: {v[5]/rows:7.0f} (rest: the peer's relay)")
print(f"per row: ring waits {v[0]/rows:7.0f}  tempty wait {v[1]/rows:7.0f}  token wait {v[2]/rows:7.0f}  issue {v[3]/rows:7.0f}  commits+handover {v[4]/rows:7.0f} cycles")
print(f"producer of CTA 0: {v[9]} rows, {v[10]/max(v[9],1):7.0f} cycles per row in total, of which {v[8]/max(v[9],1):7.0f} waiting for a free ring slot")
eng = m.engine_for(dev, 576, 752)
eng.set_profiling(True); m(x); torch.cuda.synchronize()
print("kernels:", " ".join(l for l, _ in eng.kernel_times()))
