"""Drop-in for the reference's deprecated ``model/model_conv5.py`` (:22-68 ctor, :114-151 un-fused
eval forward, :157-162 presets): float ``[B,3,H,W]`` in and out; five bias-free 3x3 convs + BatchNorm,
residual adds *before* the ReLU of layers 2 and 4, Sigmoid output.  (The reference's own
``fuse_layers`` half-mutates that model and changes its result -- SURVEY.md section 8a11 -- so the
un-fused forward is the contract.)"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import activations
from .descriptor import LayerSpec, fold_batchnorm
from .fused_module import FusedEnhancer


class Model(FusedEnhancer):
    _head = L.HEAD_PLAIN
    _tail = L.TAIL_PLAIN

    def __init__(self, initial_out_channels=32, mid_out_channels=64, final_out_channels=3, kernel_size=3):
        super().__init__()
        if kernel_size % 2 == 0:
            raise ValueError("kernel_size must be odd for symmetric padding")
        if kernel_size != 3 or final_out_channels != 3:
            raise ValueError("the fused engine implements kernel_size=3, final_out_channels=3 (both presets)")
        a, b = initial_out_channels, mid_out_channels
        chans = [(3, a), (a, a), (a, b), (b, b), (b, final_out_channels)]
        for i, (ci, co) in enumerate(chans, start=1):
            setattr(self, f"conv{i}", nn.Conv2d(ci, co, 3, 1, 1, bias=False))
            setattr(self, f"bn{i}", nn.BatchNorm2d(co))
        for i in range(1, 5):
            setattr(self, f"act{i}", activations.ReLU())
        self.act5 = activations.Sigmoid()
        self.eval()

    def fuse_layers(self):
        """BatchNorm is folded when the engine is built; the layer order is never altered."""
        return None

    def _layer_specs(self):
        fold = lambda i: fold_batchnorm(getattr(self, f"conv{i}").weight, getattr(self, f"bn{i}"))
        (w1, b1), (w2, b2), (w3, b3), (w4, b4), (w5, b5) = (fold(i) for i in range(1, 6))
        return [
            LayerSpec(w1, b1, src0=0, cin0=3, pre=[self.act1]),
            LayerSpec(w2, b2, src0=1, cin0=w2.shape[1], skip_src=1, post=[self.act2]),
            LayerSpec(w3, b3, src0=2, cin0=w3.shape[1], pre=[self.act3]),
            LayerSpec(w4, b4, src0=3, cin0=w4.shape[1], skip_src=3, post=[self.act4]),
            LayerSpec(w5, b5, src0=4, cin0=w5.shape[1], pre=[self.act5]),
        ]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._forward_float(x, 3, L.FMT_F32_NCHW3)


def get_model(name: str = "lightweight"):
    if name == "lightweight":
        return Model(initial_out_channels=32, mid_out_channels=64)
    if name == "heavyweight":
        return Model(initial_out_channels=64, mid_out_channels=128)
    return None
