"""Input side of the hot path on the GPU: colour-depth quantisation + pixel-mode replication of the reference's dataset
generator (dataset_generator/quantize.py:464-473, 512-521; util.py:318-350) and the synthetic RGB444 framebuffer
stream the benchmarks run on, generated on the device.  Thin wrappers over the C ABI; no CPU path."""
from __future__ import annotations

import torch

from . import _lib as L

COLOR_SPACES = {"RGB888": 0, "RGB444": 1, "RGB555": 2, "RGB565": 3, "RGB666": 4}
# resolution style -> (rows replicated sy, columns replicated sx): util.py:335-348
RESOLUTION_STYLES = {"lores": (2, 2), "lores_laced": (1, 2), "hires": (2, 1), "hires_laced": (1, 1)}
STYLE_OF_FRAME = ("lores", "lores_laced", "hires", "hires_laced")     # synthetic stream: frame g uses style g & 3


def _check(rc: int, what: str):
    if rc == L.ERR_INVALID:
        raise ValueError(f"{what}: invalid argument")
    if rc != 0:
        raise L.EngineError(rc, f"{what} failed")


def quantize_frames(img: torch.Tensor, color_space: str = "RGB444", style: str = "hires_laced",
                    expand17: bool = False) -> torch.Tensor:
    """``img``: CUDA uint8 ``[B,h,w,3|4]`` (the image at the style's source resolution, util.py:284-316) ->
    uint8 RGBA ``[B,h*sy,w*sx,4]`` quantised onto the colour grid and replicated to display resolution."""
    if color_space not in COLOR_SPACES:
        raise ValueError(f"Invalid color_space '{color_space}'")                       # quantize.py:477
    if style not in RESOLUTION_STYLES:
        raise ValueError(f"Unknown resolution style '{style}'")
    if not img.is_cuda or img.dtype != torch.uint8 or img.dim() != 4 or img.shape[3] not in (3, 4):
        raise ValueError("quantize_frames needs a CUDA uint8 [B,h,w,3|4] tensor")
    img = img.contiguous()
    sy, sx = RESOLUTION_STYLES[style]
    B, h, w, c = img.shape
    out = torch.empty((B, h * sy, w * sx, 4), dtype=torch.uint8, device=img.device)
    if B:
        with torch.cuda.device(img.device):
            _check(L.load().fsuae_quantize_frames(img.data_ptr(), out.data_ptr(), B, h, w, c, COLOR_SPACES[color_space], sy, sx,
                                                  1 if expand17 else 0, torch.cuda.current_stream(img.device).cuda_stream),
                   "fsuae_quantize_frames")
    return out


DITHER_METHODS = {"none": 0, "checkerboard": 1, "bayer2x2": 2, "bayer4x4": 3, "bayer8x8": 4}


def dither_frames(img: torch.Tensor, palette: torch.Tensor, method: str = "checkerboard") -> torch.Tensor:
    """Palette dithers of the reference's dataset generator (quantize.py:137-331; 'none' = nearest palette colour, :529-537).
    ``img``: CUDA uint8 ``[B,h,w,3|4]``; ``palette``: uint8 ``[N,3]`` (N <= 4096, any device) -> uint8 RGBA ``[B,h,w,4]``.
    Error-diffusion methods are sequential CPU algorithms and stay in the reference's offline generator."""
    if method not in DITHER_METHODS:
        raise ValueError(f"dithering_method must be one of {sorted(DITHER_METHODS)}")       # quantize.py:434-436
    if not img.is_cuda or img.dtype != torch.uint8 or img.dim() != 4 or img.shape[3] not in (3, 4):
        raise ValueError("dither_frames needs a CUDA uint8 [B,h,w,3|4] tensor")
    if palette.dtype != torch.uint8 or palette.dim() != 2 or palette.shape[1] != 3 or palette.shape[0] > 4096:
        raise ValueError("palette must be uint8 [N,3] with N <= 4096")
    img = img.contiguous()
    pal = palette.to(img.device).contiguous()
    B, h, w, c = img.shape
    out = torch.empty((B, h, w, 4), dtype=torch.uint8, device=img.device)
    if B:
        with torch.cuda.device(img.device):
            _check(L.load().fsuae_dither_frames(img.data_ptr(), out.data_ptr(), B, h, w, c, pal.data_ptr() if pal.numel() else None,
                                                pal.shape[0], DITHER_METHODS[method],
                                                torch.cuda.current_stream(img.device).cuda_stream), "fsuae_dither_frames")
    return out


def synth_rgb444_frames(n_frames: int, height: int = 576, width: int = 752, seed: int = 0, first_frame: int = 0,
                        expand17: bool = True, device=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """``n_frames`` synthetic RGB444 framebuffers ``[n,H,W,4]`` uint8 on the GPU; frame ``first_frame + i`` uses pixel mode
    ``(first_frame + i) & 3``.  Deterministic in ``(seed, frame index)``: any sharding of a stream gives the same frames."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if out is None:
        out = torch.empty((n_frames, height, width, 4), dtype=torch.uint8, device=device)
    elif not (out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and tuple(out.shape) == (n_frames, height, width, 4)):
        raise ValueError("out must be a contiguous CUDA uint8 [n,H,W,4] tensor")
    if n_frames:
        with torch.cuda.device(out.device):
            _check(L.load().fsuae_synth_rgb444_frames(out.data_ptr(), n_frames, height, width, seed & (2 ** 64 - 1), first_frame,
                                                      1 if expand17 else 0, torch.cuda.current_stream(out.device).cuda_stream),
                   "fsuae_synth_rgb444_frames")
    return out
