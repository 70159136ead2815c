"""Drop-in for the reference's ``model/model_pix_shuffle.py``: same constructor arguments, same
``state_dict`` keys, same ``forward(x)`` contract (float ``[B,3,H,W]`` linear-light RGB in and
out, H and W even), same ``get_model`` presets -- executed by the fused B200 engine.

Network (reference model_pix_shuffle.py:227-298): PixelUnshuffle(2) -> conv1..conv7 (3x3, zero
pad 1, bias) with 2-4 activation slots per layer, residual adds after conv2 / conv4, long-skip
concat into conv6, PixelShuffle(2), + input, ReLU.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import activations
from .descriptor import LayerSpec
from .fused_module import FusedEnhancer

_DEFAULT_ACTS = {  # constructor defaults, reference model_pix_shuffle.py:20-68
    (1, 1): "identity", (1, 2): "relu",
    (2, 1): "mish", (2, 2): "biased_relu", (2, 3): "tanh", (2, 4): "relu6",
    (3, 1): "identity", (3, 2): "identity",
    (4, 1): "telu", (4, 2): "leaky_relu", (4, 3): "tanh", (4, 4): "identity",
    (5, 1): "identity", (5, 2): "identity",
    (6, 1): "mish", (6, 2): "prelu",
    (7, 1): "sinlu", (7, 2): "prelu",
}


class Model(FusedEnhancer):
    _head = L.HEAD_UNSHUFFLE2
    _tail = L.TAIL_SHUFFLE2_RESIDUAL_RELU

    def __init__(self, verbose: bool = False, **kwargs):
        """Accepts the reference's keyword arguments: ``layer{1..6}_out_channels`` (default 36),
        ``layer{1..7}_kernel_size`` (default 3; 1, 3, 5 or 7 here), ``layer{L}_act{K}`` and ``layer{L}_act{K}_params``."""
        super().__init__()
        self.verbose = verbose
        ch = [int(kwargs.pop(f"layer{i}_out_channels", 36)) for i in range(1, 7)]
        ks = [int(kwargs.pop(f"layer{i}_kernel_size", 3)) for i in range(1, 8)]
        for k in ks:
            if k % 2 == 0:
                raise ValueError("kernel_size must be odd for symmetric padding")      # reference :81-83
        if any(k > 7 for k in ks):
            raise ValueError("the fused engine implements kernel sizes 1, 3, 5 and 7")
        self.channels = tuple(ch)
        self.kernel_sizes = tuple(ks)
        cin = [12, ch[0], ch[1], ch[2], ch[3], ch[0] + ch[4], ch[5]]
        cout = ch + [12]
        self.pixel_unshuffle = nn.PixelUnshuffle(2)   # structure marker only (no parameters)
        for i in range(7):
            setattr(self, f"conv{i + 1}", nn.Conv2d(cin[i], cout[i], ks[i], stride=1, padding=(ks[i] - 1) // 2, bias=True))
        for (layer, idx), default in _DEFAULT_ACTS.items():
            name = kwargs.pop(f"layer{layer}_act{idx}", default)
            params = kwargs.pop(f"layer{layer}_act{idx}_params", None)
            setattr(self, f"l{layer}_act{idx}", activations.get_activation(name, params=params))
        # 1x1 projections of the short skips when the channel counts differ (reference :126-128, :143-145; same keys)
        self.skip1_proj_conv = nn.Conv2d(ch[0], ch[1], 1, stride=1, padding=0, bias=False) if ch[0] != ch[1] else None
        self.skip2_proj_conv = nn.Conv2d(ch[2], ch[3], 1, stride=1, padding=0, bias=False) if ch[2] != ch[3] else None
        self.pixel_shuffle = nn.PixelShuffle(2)
        if kwargs:
            raise TypeError(f"unexpected arguments: {sorted(kwargs)}")

    def fuse_layers(self):
        """No-op, as in the reference (its fuse_modules call raises and is swallowed):
        the whole network is already one fused pass."""
        return None

    def _layer_specs(self):
        c = lambda i: getattr(self, f"conv{i}")
        a = lambda l, k: getattr(self, f"l{l}_act{k}")
        ch = self.channels
        specs = []

        def add(spec):       # -> buffer id of the layer's output (0 = unshuffled input, i = output of specs[i-1])
            specs.append(spec)
            return len(specs)

        def projected(proj, src, cin):
            """A 1x1 skip projection runs as a layer of its own: centre tap of a 3x3 kernel, no bias, no activation."""
            if proj is None:
                return src
            w = F.pad(proj.weight, (1, 1, 1, 1))
            return add(LayerSpec(w, None, src0=src, cin0=cin))

        b1 = add(LayerSpec(c(1).weight, c(1).bias, src0=0, cin0=12, pre=[a(1, 1), a(1, 2)]))
        s1 = projected(self.skip1_proj_conv, b1, ch[0])
        b2 = add(LayerSpec(c(2).weight, c(2).bias, src0=b1, cin0=ch[0], skip_src=s1,
                           pre=[a(2, 1), a(2, 2)], post=[a(2, 3), a(2, 4)]))
        b3 = add(LayerSpec(c(3).weight, c(3).bias, src0=b2, cin0=ch[1], pre=[a(3, 1), a(3, 2)]))
        s2 = projected(self.skip2_proj_conv, b3, ch[2])
        b4 = add(LayerSpec(c(4).weight, c(4).bias, src0=b3, cin0=ch[2], skip_src=s2,
                           pre=[a(4, 1), a(4, 2)], post=[a(4, 3), a(4, 4)]))
        b5 = add(LayerSpec(c(5).weight, c(5).bias, src0=b4, cin0=ch[3], pre=[a(5, 1), a(5, 2)]))
        b6 = add(LayerSpec(c(6).weight, c(6).bias, src0=b1, cin0=ch[0], src1=b5, cin1=ch[4], pre=[a(6, 1), a(6, 2)]))
        add(LayerSpec(c(7).weight, c(7).bias, src0=b6, cin0=ch[5], pre=[a(7, 1), a(7, 2)]))
        return specs

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: ``[B,3,H,W]`` float (fp32; fp16/bf16 are converted), H and W even."""
        if x.dim() == 4 and (x.shape[2] % 2 or x.shape[3] % 2):
            raise ValueError("PixelUnshuffle(2) needs even height and width")
        return self._forward_float(x, 3, L.FMT_F32_NCHW3)


def get_model(name: str = "lightweight"):
    """Presets of reference model_pix_shuffle.py:304-314."""
    if name == "lightweight":
        return Model(layer1_out_channels=36, layer2_out_channels=36, layer3_out_channels=72,
                     layer4_out_channels=72, layer5_out_channels=36, layer6_out_channels=36,
                     layer1_act1="sinlu", layer1_act2="relu6",
                     layer2_act1="telu", layer2_act2="identity", layer2_act3="sinlu",
                     layer2_act4="biased_prelu", layer2_act4_params={"num_parameters": 36},
                     layer4_act1="mish", layer4_act2="biased_prelu", layer4_act2_params={"num_parameters": 72},
                     layer4_act3="tanh", layer4_act4="relu",
                     layer6_act1="mish", layer6_act2="relu6",
                     layer7_act1="identity", layer7_act2="biased_prelu", layer7_act2_params={"num_parameters": 1})
    if name == "heavyweight":
        return Model(layer1_out_channels=36, layer2_out_channels=36, layer3_out_channels=108,
                     layer4_out_channels=108, layer5_out_channels=36, layer6_out_channels=36)
    return None
