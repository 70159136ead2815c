#!/usr/bin/env python3
"""Device-resident throughput of every model family / preset / build at 752x576 (BASELINE.json configs 3-4 are
parity cases, this table is informational).  CUDA events, 3 warm-ups, inputs rotate over two batches."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs_uae_image_enhancer_project_b200 import model_pix_shuffle, model_conv3, model_conv5, _lib
dev = torch.device("cuda", 0)
H, W = 576, 752
GF = {("pix_shuffle", "lightweight"): 29.472, ("pix_shuffle", "heavyweight"): 47.155, ("conv3", "lightweight"): 18.213,
      ("conv3", "heavyweight"): 393.704, ("conv5", "lightweight"): 58.132, ("conv5", "heavyweight"): 228.039}
rows = []
for fam, mod, batch in (("pix_shuffle", model_pix_shuffle, 32), ("conv3", model_conv3, 16), ("conv5", model_conv5, 32)):
    for preset in ("lightweight", "heavyweight"):
        for prec in ("bf16", "fp32"):
            m = mod.get_model(preset).to(dev).set_precision(prec)
            m.chunk_frames = batch
            b = batch if prec == "bf16" else 4
            if fam == "conv3":
                xs = [torch.randint(0, 256, (b, 4, H, W), dtype=torch.uint8, device=dev) for _ in range(2)]
            else:
                xs = [torch.rand(b, 3, H, W, device=dev) for _ in range(2)]
            try:
                for i in range(3):
                    m(xs[i & 1])
            except _lib.EngineError as exc:
                rows.append((fam, preset, prec, None, str(exc)[:60]))
                continue
            torch.cuda.synchronize()
            steps = 6 if prec == "bf16" else 2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                m(xs[i & 1])
            e1.record(); torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / (steps * b)
            rows.append((fam, preset, prec, us, f"{1e6 / us:9.0f} fps  {GF[(fam, preset)] / us * 1e3:7.0f} TFLOP/s"))
            m.close()
for r in rows:
    print(f"{r[0]:12s} {r[1]:12s} {r[2]:5s} " + (f"{r[3]:9.1f} us/frame  {r[4]}" if r[3] else f"unsupported: {r[4]}"))
