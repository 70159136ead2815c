#!/usr/bin/env python3
"""Debug aid: the single fused pass (csrc/mega.cu) against the layer-by-layer kernels of the same build, same weights and
frames: largest difference of the uint8 framebuffers / float outputs, and device time per frame of both.

    python tools/mega_check.py [frames] [reps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs_uae_image_enhancer_project_b200 import model_pix_shuffle  # noqa: E402

dev = torch.device("cuda", 0)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.manual_seed(3)
ref = model_pix_shuffle.get_model("lightweight")
sd = ref.state_dict()
g = torch.Generator().manual_seed(1)
fb = (torch.randint(0, 16, (b, 576, 752, 4), generator=g, dtype=torch.uint8) * 17).to(dev)
fb[..., 3] = 255
xf = torch.rand(b, 3, 576, 752, generator=g).to(dev)


def build(no_mega):
    if no_mega:
        os.environ["FSUAE_NO_MEGA"] = "1"
    else:
        os.environ.pop("FSUAE_NO_MEGA", None)
        os.environ.setdefault("FSUAE_MEGA_MIN_FRAMES", "2")      # float frames too (by default they stay on the layer kernels)
    m = model_pix_shuffle.get_model("lightweight")
    m.load_state_dict(sd)
    m = m.to(dev).set_precision("bf16")
    m.chunk_frames = b
    return m


def timed(m, fn, x):
    for _ in range(3):
        fn(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        y = fn(x)
    e1.record()
    torch.cuda.synchronize()
    return y, e0.elapsed_time(e1) * 1e3 / (reps * b)


res = {}
for name, no_mega in (("layers", True), ("fused_pass", False)):
    m = build(no_mega)
    yu, tu = timed(m, m.forward_framebuffer, fb)
    yf, tf = timed(m, m.forward, xf)
    eng = m.engine_for(dev, 576, 752)
    print(f"{name:10s}: u8 {tu:7.2f} us/frame   f32 {tf:7.2f} us/frame   launches/pass {eng.last_launch_count}  [{eng.variant}]")
    res[name] = (yu.cpu(), yf.cpu())
du = (res["layers"][0].int() - res["fused_pass"][0].int()).abs()
df = (res["layers"][1] - res["fused_pass"][1]).abs()
print(f"fused pass vs layers: u8 max LSB diff {du.max().item()}  differing {100.0 * (du > 0).float().mean().item():.4f} %   f32 max|d| {df.max().item():.3e}")
bad = du.max().item() > 2 or df.max().item() > 2e-2
print("MEGA_CHECK", "FAIL" if bad else "OK")
sys.exit(1 if bad else 0)
