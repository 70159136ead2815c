#!/usr/bin/env python3
"""Debug aid: does host <-> device copy traffic slow the device-resident pass down (and the pass the copies)?

    python tools/copy_interference.py [frames]
"""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fs_uae_image_enhancer_project_b200 import model_pix_shuffle  # noqa: E402

dev = torch.device("cuda", 0)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = model_pix_shuffle.get_model("lightweight").to(dev).set_precision("bf16")
m.chunk_frames = b
x = (torch.randint(0, 16, (b, 576, 752, 4), dtype=torch.uint8) * 17).to(dev)
h_in = torch.empty_like(x, device="cpu").pin_memory()
h_out = torch.empty_like(x, device="cpu").pin_memory()
d_in, d_out = torch.empty_like(x), torch.empty_like(x)
s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
stop = False


def copier():
    while not stop:
        with torch.cuda.stream(s_up):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_down):
            h_out.copy_(d_out, non_blocking=True)
        s_up.synchronize()
        s_down.synchronize()


def timed(reps=40):
    for _ in range(5):
        m.forward_framebuffer(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        m.forward_framebuffer(x)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * b)


def copy_rate(seconds=0.5):
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        with torch.cuda.stream(s_up):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_down):
            h_out.copy_(d_out, non_blocking=True)
        s_up.synchronize()
        s_down.synchronize()
        n += 1
    return n * b / (time.perf_counter() - t0)


print(f"{b}-frame passes, {m.engine_for(dev, 576, 752).variant}")
print(f"pass alone          : {timed():6.2f} us/frame")
print(f"copies alone        : {copy_rate():8.0f} frames/s both ways")
th = threading.Thread(target=copier)
th.start()
time.sleep(0.05)
print(f"pass under copies   : {timed():6.2f} us/frame")
stop = True
th.join()


def runner():
    while not stop2:
        for _ in range(4):
            m.forward_framebuffer(x)
        torch.cuda.current_stream().synchronize()


stop2 = False
th = threading.Thread(target=runner)
th.start()
time.sleep(0.05)
print(f"copies under passes : {copy_rate():8.0f} frames/s both ways")
stop2 = True
th.join()
