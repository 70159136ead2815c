"""Drop-in for the reference's ``model/residual_feature_block.py`` (:5-55), the bottleneck building block of
``model/model_residual_unet.py``: 1x1 conv (in -> mid) -> k x k conv (mid -> mid) -> act1, act2 -> 1x1 conv (mid -> out) -> act3
-> + identity (through a biased 1x1 projection when in != out) -> act4.  Same constructor arguments, same ``state_dict`` keys
(``conv1..3``, ``act1..4``, ``proj_conv``), same ``forward(x)``: float ``[B,in,H,W]`` -> float ``[B,out,H,W]``.

On the engine the block is three (four with the projection) layers of the network descriptor: the 1x1 convolutions run as
the centre tap of a 3x3 kernel, the residual comes from the block's input (or the projection's buffer), kernel sizes 5 and 7
use the window decomposition of the K-streamed tensor-core kernel.  Feature maps enter and leave as float NCHW tensors
(``FSUAE_HEAD_FEATURES`` / ``FSUAE_TAIL_FEATURES``)."""
from __future__ import annotations

import copy

import torch
import torch.nn as nn

from . import _lib as L
from . import activations
from .descriptor import LayerSpec
from .fused_module import FusedEnhancer

_DEFAULT_ACTS = {"act1": "identity", "act1_params": None, "act2": "relu", "act2_params": None,
                 "act3": "identity", "act3_params": None, "act4": "relu", "act4_params": None}


class ResidualFeatureBlock(FusedEnhancer):
    _head = L.HEAD_FEATURES
    _tail = L.TAIL_FEATURES

    def __init__(self, in_channels, mid_channels, out_channels, kernel_size, acts=None):
        super().__init__()
        if kernel_size % 2 == 0:
            raise ValueError("kernel_size must be odd for symmetric padding")          # reference :17-18
        if kernel_size > 7:
            raise ValueError("the fused engine implements kernel sizes 1, 3, 5 and 7")
        acts = copy.deepcopy(acts if acts is not None else _DEFAULT_ACTS)
        self._in_channels, self.mid_channels, self.out_channels = int(in_channels), int(mid_channels), int(out_channels)
        self.conv1 = nn.Conv2d(in_channels, mid_channels, 1, stride=1, padding=0, bias=True)
        self.conv2 = nn.Conv2d(mid_channels, mid_channels, kernel_size, stride=1, padding=(kernel_size - 1) // 2, bias=True)
        self.conv3 = nn.Conv2d(mid_channels, out_channels, 1, stride=1, padding=0, bias=True)
        for key, ch in zip(("act1", "act2", "act3", "act4"), (mid_channels, mid_channels, out_channels, out_channels)):
            params = acts.get(f"{key}_params")
            if isinstance(params, dict):                                                # reference :24-35
                if params.get("num_parameters") == "global":
                    params["num_parameters"] = 1
                elif params.get("num_parameters") == "channel":
                    params["num_parameters"] = ch
            setattr(self, key, activations.get_activation(acts[key], params=params))
        self.proj_conv = nn.Conv2d(in_channels, out_channels, 1, stride=1, padding=0, bias=True) if in_channels != out_channels else None

    def _layer_specs(self):
        specs = []

        def add(spec):
            specs.append(spec)
            return len(specs)

        b1 = add(LayerSpec(self.conv1.weight, self.conv1.bias, src0=0, cin0=self._in_channels))
        b2 = add(LayerSpec(self.conv2.weight, self.conv2.bias, src0=b1, cin0=self.mid_channels, pre=[self.act1, self.act2]))
        ident = 0
        if self.proj_conv is not None:
            ident = add(LayerSpec(self.proj_conv.weight, self.proj_conv.bias, src0=0, cin0=self._in_channels))
        add(LayerSpec(self.conv3.weight, self.conv3.bias, src0=b2, cin0=self.mid_channels, skip_src=ident,
                      pre=[self.act3], post=[self.act4]))
        return specs

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._forward_features(x, self.out_channels)
