#!/usr/bin/env python3
"""PCIe ceiling of the box: pinned host <-> device copies alone and both directions at once (the end-to-end bound of
fsuae_engine_run_host: 1.73 MB in + 1.73 MB out per frame)."""
import torch, time
dev = torch.device("cuda", 0)
for mb in (16, 111):   # 111 MB = 64 frames of uint8 RGBA
    n = mb << 20
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(h2d, d2h, reps=20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return reps * n / (time.perf_counter() - t0) / 1e9
    run(True, True, 3)
    print(f"{mb:4d} MB: H2D alone {run(True, False):6.1f} GB/s   D2H alone {run(False, True):6.1f} GB/s   "
          f"both at once {run(True, True):6.1f} GB/s per direction")
