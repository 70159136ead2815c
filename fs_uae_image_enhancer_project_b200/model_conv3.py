"""Drop-in for the reference's deprecated ``model/model_conv3.py`` (:20-55 ctor, :102-155 forward,
:206-211 presets): uint8 planar RGBA ``[B,4,H,W]`` in, float ``[B,4,H,W]`` out (RGB x255, unclipped,
alpha 255.0); three bias-free 3x3 convs each followed by BatchNorm (folded on the host), ReLU after
the first two."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import activations
from .descriptor import LayerSpec, fold_batchnorm
from .fused_module import FusedEnhancer


class Model(FusedEnhancer):
    _head = L.HEAD_PLAIN
    _tail = L.TAIL_SCALE255_ALPHA

    def __init__(self, initial_out_channels=32, mid_out_channels=64, final_out_channels=3, kernel_size=3):
        super().__init__()
        if kernel_size % 2 == 0:
            raise ValueError("kernel_size must be odd for symmetric padding")
        if kernel_size != 3 or final_out_channels != 3:
            raise ValueError("the fused engine implements kernel_size=3, final_out_channels=3 (both presets)")
        chans = [3, initial_out_channels, mid_out_channels, final_out_channels]
        for i in range(3):
            setattr(self, f"conv{i + 1}", nn.Conv2d(chans[i], chans[i + 1], 3, 1, 1, bias=False))
            setattr(self, f"bn{i + 1}", nn.BatchNorm2d(chans[i + 1]))
        self.act1 = activations.ReLU()
        self.act2 = activations.ReLU()
        self.eval()

    def fuse_layers(self):
        """BatchNorm is always folded into the convs when the engine is built; nothing to do."""
        return None

    def _layer_specs(self):
        specs = []
        for i in range(1, 4):
            w, b = fold_batchnorm(getattr(self, f"conv{i}").weight, getattr(self, f"bn{i}"))
            pre = [getattr(self, f"act{i}")] if i < 3 else []
            specs.append(LayerSpec(w, b, src0=i - 1, cin0=w.shape[1], pre=pre))
        return specs

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[1] != 4:
            raise ValueError("Input tensor must be uint8 with 4 channels (RGBA)")
        self._require_cuda(x)
        x = x.contiguous()
        B, _, H, W = x.shape
        out = torch.empty((B, 4, H, W), dtype=torch.float32, device=x.device)
        if B > 0:
            self.engine_for(x.device, H, W).enqueue(x, out, B, L.FMT_U8_NCHW4, L.FMT_F32_NCHW4)
        p = next(self.parameters())
        return out if p.dtype == torch.float32 else out.to(p.dtype)


def get_model(name: str = "lightweight"):
    if name == "lightweight":
        return Model(initial_out_channels=32, mid_out_channels=64)
    if name == "heavyweight":
        return Model(initial_out_channels=192, mid_out_channels=256)
    return None
