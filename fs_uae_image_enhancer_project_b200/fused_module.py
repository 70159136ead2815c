"""Base class of the drop-in models: an ``nn.Module`` that owns the reference's parameters and runs
its forward as one call into the B200 engine (C ABI, include/fsuae_enhancer.h)."""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from .descriptor import LayerSpec, build_descriptor
from .engine import Engine


class FusedEnhancer(nn.Module):
    """Subclasses provide ``_layer_specs()``, ``_head`` and ``_tail``.

    * fp32 parameters -> the fp32 FMA build; after ``.half()`` -> the fp16 tensor-core build (what the
      reference deploys, torch2onnx.py:58); after ``.bfloat16()`` -> the bf16 tensor-core build
      (``precision`` can also be forced with ``set_precision``).  Any other parameter dtype raises.
    * ``load_state_dict`` accepts genuine reference checkpoints: the ``perceptual_criterion.*`` entries
      (the VGG16 of the training loss, saved by train.py:236/246 with every ``state_dict``) are dropped.
    * Engines are cached per (device, H, W, precision) and rebuilt when a parameter changes
      (``load_state_dict``, optimiser step, ``.to``).
    * CUDA tensors only.  There is no CPU path: a CPU tensor raises.
    """
    _head = L.HEAD_PLAIN
    _tail = L.TAIL_PLAIN
    _in_channels = 0          # HEAD_FEATURES only
    chunk_frames = 8

    def __init__(self):
        super().__init__()
        self._engines: Dict[Tuple, Engine] = {}
        self._param_stamp = None
        self._forced_precision: Optional[int] = None

    # -- to be provided by subclasses ---------------------------------------------------------
    def _layer_specs(self) -> Sequence[LayerSpec]:
        raise NotImplementedError

    # -- engine management --------------------------------------------------------------------
    def set_precision(self, precision: Optional[str]):
        """Force 'fp32', 'fp16' or 'bf16' regardless of the parameter dtype (None: follow the parameters)."""
        self._forced_precision = {None: None, "fp32": L.PREC_FP32, "bf16": L.PREC_BF16, "fp16": L.PREC_FP16}[precision]
        return self

    def _precision(self) -> int:
        if self._forced_precision is not None:
            return self._forced_precision
        p = next(self.parameters())
        try:
            return {torch.float32: L.PREC_FP32, torch.float16: L.PREC_FP16, torch.bfloat16: L.PREC_BF16}[p.dtype]
        except KeyError:
            raise TypeError(f"parameters are {p.dtype}: the engine has fp32, fp16 and bf16 builds only") from None

    # -- checkpoints ----------------------------------------------------------------------------
    _IGNORED_STATE_PREFIXES = ("perceptual_criterion.",)

    def load_state_dict(self, state_dict, strict: bool = True, **kwargs):
        """Reference checkpoints carry the training loss's VGG16 (`self.perceptual_criterion = PerceptualLoss(...)`,
        model_pix_shuffle.py:172-182; saved by train.py:236/246): those keys are not part of the network and are
        dropped, so `m.load_state_dict(torch.load('best.pth'))` works with the default strict=True."""
        kept = {k: v for k, v in state_dict.items() if not k.startswith(self._IGNORED_STATE_PREFIXES)}
        return super().load_state_dict(kept, strict=strict, **kwargs)

    def _stamp(self):
        return tuple((id(p), p._version, p.dtype, p.device) for p in list(self.parameters()) + list(self.buffers()))

    def engine_for(self, device: torch.device, height: int, width: int) -> Engine:
        stamp = self._stamp()
        if stamp != self._param_stamp:
            for e in self._engines.values():
                e.close()
            self._engines.clear()
            self._param_stamp = stamp
        key = (device.index if device.index is not None else torch.cuda.current_device(), height, width,
               self._precision(), self.chunk_frames)
        eng = self._engines.get(key)
        if eng is None:
            desc, blob = build_descriptor(self._layer_specs(), self._head, self._tail, self._in_channels)
            eng = Engine(desc, blob, key[0], key[3], height, width, self.chunk_frames)
            self._engines[key] = eng
        return eng

    def close(self):
        for e in self._engines.values():
            e.close()
        self._engines.clear()

    # -- forward variants ---------------------------------------------------------------------
    def _require_cuda(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("this model runs only on a CUDA (sm_100a) device: the fused engine has no CPU path; "
                               "move the input with .cuda() or use run_host() for pinned host buffers")

    def _forward_float(self, x: torch.Tensor, out_channels: int, out_fmt: int) -> torch.Tensor:
        self._require_cuda(x)
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected a [B,3,H,W] tensor, got {tuple(x.shape)}")
        in_dtype = x.dtype
        xf = x.contiguous() if x.dtype == torch.float32 else x.float().contiguous()
        B, _, H, W = xf.shape
        out = torch.empty((B, out_channels, H, W), dtype=torch.float32, device=x.device)
        if B == 0:
            return out.to(in_dtype)
        self.engine_for(x.device, H, W).enqueue(xf, out, B, L.FMT_F32_NCHW3, out_fmt)
        return out if in_dtype == torch.float32 else out.to(in_dtype)

    def _forward_features(self, x: torch.Tensor, out_channels: int) -> torch.Tensor:
        """Feature-map networks (HEAD_FEATURES / TAIL_FEATURES): float ``[B,C,H,W]`` in, float ``[B,out_channels,H,W]`` out."""
        self._require_cuda(x)
        if x.dim() != 4 or x.shape[1] != self._in_channels:
            raise ValueError(f"expected a [B,{self._in_channels},H,W] tensor, got {tuple(x.shape)}")
        in_dtype = x.dtype
        xf = x.contiguous() if x.dtype == torch.float32 else x.float().contiguous()
        B, _, H, W = xf.shape
        out = torch.empty((B, out_channels, H, W), dtype=torch.float32, device=x.device)
        if B:
            self.engine_for(x.device, H, W).enqueue(xf, out, B, L.FMT_F32_NCHW, L.FMT_F32_NCHW)
        return out if in_dtype == torch.float32 else out.to(in_dtype)

    def forward_framebuffer(self, rgba: torch.Tensor, gamma: bool = True, crop16: bool = False) -> torch.Tensor:
        """Deployed contract: uint8 ``[B,H,W,4]`` RGBA in -> uint8 ``[B,H,W,4]`` out, alpha 255
        (reference convertion_tools/torch2onnx.py:184-768).  ``gamma``: /255, **2.2 in front and
        **(1/2.2), x255, clip, truncate behind; ``crop16``: newer exporter's 16-column crop/pad."""
        self._require_cuda(rgba)
        if rgba.dtype != torch.uint8 or rgba.dim() != 4 or rgba.shape[3] != 4:
            raise ValueError("Input tensor must be uint8 [B,H,W,4] (RGBA, chunky)")
        rgba = rgba.contiguous()
        B, H, W, _ = rgba.shape
        out = torch.empty_like(rgba)
        if B == 0:
            return out
        flags = (L.FLAG_GAMMA_IN | L.FLAG_GAMMA_OUT if gamma else 0) | (L.FLAG_CROP16 if crop16 else 0)
        self.engine_for(rgba.device, H, W).enqueue(rgba, out, B, L.FMT_U8_NHWC4, L.FMT_U8_NHWC4, flags)
        return out

    def run_host(self, rgba_host: torch.Tensor, out_host: Optional[torch.Tensor] = None, gamma: bool = True,
                 crop16: bool = False, device: int = 0) -> torch.Tensor:
        """End-to-end framebuffer call with HOST uint8 ``[B,H,W,4]`` buffers (pinned recommended):
        upload, fused forward, copy back (reference README.md:21-24)."""
        if rgba_host.is_cuda or rgba_host.dtype != torch.uint8 or rgba_host.dim() != 4 or rgba_host.shape[3] != 4:
            raise ValueError("run_host needs a CPU uint8 [B,H,W,4] tensor")
        rgba_host = rgba_host.contiguous()
        B, H, W, _ = rgba_host.shape
        if out_host is None:
            out_host = torch.empty_like(rgba_host, pin_memory=rgba_host.is_pinned())
        flags = (L.FLAG_GAMMA_IN | L.FLAG_GAMMA_OUT if gamma else 0) | (L.FLAG_CROP16 if crop16 else 0)
        dev = torch.device("cuda", device)
        self.engine_for(dev, H, W).run_host(rgba_host, out_host, B, L.FMT_U8_NHWC4, L.FMT_U8_NHWC4, flags)
        return out_host
