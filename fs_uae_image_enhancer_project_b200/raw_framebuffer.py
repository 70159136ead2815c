#!/usr/bin/env python3
"""Raw-framebuffer tool: the B200 counterpart of the reference's
``convertion_tools/convert_raw_to_png_using_final_model.py`` (:10-37 raw loader, :60-93 inference).

    python -m fs_uae_image_enhancer_project_b200.raw_framebuffer WEIGHTS IN.raw [OUT.raw|OUT.png]
        [--width 752 --height 576] [--precision bf16|fp32] [--crop16] [--onnx]

``IN.raw`` holds ``width*height*4`` bytes, RGBA, row-major, one or more frames back to back.
``WEIGHTS`` is a ``state_dict`` saved with ``torch.save`` for ``model_pix_shuffle.get_model('lightweight')``
or, with ``--onnx``, the reference's shipped ``pix_shuffle.onnx`` (trained weights are read out of it).
The frames go through the fused engine's uint8 framebuffer contract; errors exit with status 1 like
the reference tool.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch


def load_raw_rgba(path: str, width: int = 752, height: int = 576) -> torch.Tensor:
    """-> uint8 [N,H,W,4]; the file size must be a whole number of frames."""
    data = np.fromfile(path, dtype=np.uint8)
    frame = width * height * 4
    if data.size == 0 or data.size % frame:
        raise ValueError(f"Expected raw file of a multiple of {frame} bytes ({width}x{height} RGBA), but got {data.size} bytes.")
    return torch.from_numpy(data.reshape(-1, height, width, 4))


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="Run raw RGBA framebuffers through the fused B200 enhancer.")
    ap.add_argument("weights")
    ap.add_argument("raw_path")
    ap.add_argument("out_path", nargs="?")
    ap.add_argument("--width", type=int, default=752)
    ap.add_argument("--height", type=int, default=576)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--crop16", action="store_true", help="newer exporter contract: 16 black columns on the left")
    ap.add_argument("--onnx", action="store_true", help="WEIGHTS is the reference's pix_shuffle.onnx")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    try:
        from . import model_pix_shuffle, onnx_weights
        frames = load_raw_rgba(args.raw_path, args.width, args.height)
        sd = onnx_weights.pix_shuffle_state_dict_from_onnx(args.weights) if args.onnx else torch.load(args.weights, map_location="cpu")
        model = model_pix_shuffle.get_model("lightweight")
        model.load_state_dict(sd)
        model.set_precision(args.precision)
        out = model.run_host(frames.pin_memory() if torch.cuda.is_available() else frames, crop16=args.crop16, device=args.device)
    except Exception as exc:  # noqa: BLE001  (the reference tool prints and exits 1)
        print(f"Error: {exc}")
        return 1
    out_path = args.out_path or os.path.splitext(args.raw_path)[0] + ".png"
    if out_path.endswith(".png"):
        from PIL import Image
        Image.fromarray(out[0].numpy(), mode="RGBA").save(out_path)
    else:
        out.numpy().tofile(out_path)
    print(f"Saved output to '{out_path}'")
    return 0


if __name__ == "__main__":
    sys.exit(main())
