"""Network descriptor + parameter blob for the C ABI (include/fsuae_enhancer.h).

A descriptor is derived from a *module instance* (not from state_dict key names -- which keys exist
depends on constructor kwargs): per 3x3 conv layer its sources, skip wiring,
BatchNorm-folded fp32 weights/bias and the activation slots around the skip add.  Everything is
flattened into one float32 blob addressed by offsets.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L


@dataclass
class LayerSpec:
    weight: torch.Tensor                    # [Cout, Cin0+Cin1, k, k], k in {1, 3, 5, 7} (BN folded); 1x1 runs as the centre tap of 3x3
    bias: Optional[torch.Tensor]            # [Cout] or None
    src0: int
    cin0: int
    src1: int = -1
    cin1: int = 0
    skip_src: int = -1
    pre: Sequence = field(default_factory=list)    # activation modules applied before the skip add
    post: Sequence = field(default_factory=list)   # ... and after it


class BlobBuilder:
    def __init__(self):
        self.parts: List[np.ndarray] = []
        self.size = 0

    def add(self, t) -> int:
        a = t.detach().to(torch.float32).cpu().numpy().reshape(-1) if isinstance(t, torch.Tensor) \
            else np.asarray(t, dtype=np.float32).reshape(-1)
        off = self.size
        self.parts.append(np.ascontiguousarray(a, dtype=np.float32))
        self.size += a.size
        pad = (-self.size) % 4                       # keep every entry 16-byte aligned
        if pad:
            self.parts.append(np.zeros(pad, np.float32))
            self.size += pad
        return off

    def finish(self) -> np.ndarray:
        return np.concatenate(self.parts) if self.parts else np.zeros(0, np.float32)


def fold_batchnorm(weight: torch.Tensor, bn) -> (torch.Tensor, torch.Tensor):
    """Eval-mode BatchNorm2d folded into the preceding bias-free conv, in fp32 on the host:
    W' = W * gamma / sqrt(var + eps), b' = beta - mean * gamma / sqrt(var + eps)
    (reference model_conv3.py:41-52 / model_conv5.py:42-64, BN eps 1e-5)."""
    w = weight.detach().to(torch.float32)
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return w * scale.view(-1, 1, 1, 1), bn.bias.detach().float() - bn.running_mean.detach().float() * scale


def _fill_act(dst: L.ActDesc, module, channels: int, blob: BlobBuilder):
    name, params = module.engine_op(channels)
    dst.op = L.ACT[name]
    dst.n0 = dst.n1 = 0
    dst.p0_off = dst.p1_off = -1
    for i, p in enumerate(params):
        n = int(p.numel())
        if n != 1 and n != channels:
            # the reference would broadcast such a tensor along W or fail; neither is a network
            raise ValueError(f"activation '{name}': parameter count {n} must be 1 or the channel count {channels}")
        off = blob.add(p)
        if i == 0:
            dst.n0, dst.p0_off = n, off
        else:
            dst.n1, dst.p1_off = n, off


def build_descriptor(layers: Sequence[LayerSpec], head: int, tail: int, in_channels: int = 0):
    """-> (NetDesc, float32 blob).  Identity slots are dropped."""
    if not 1 <= len(layers) <= L.MAX_LAYERS:
        raise ValueError(f"between 1 and {L.MAX_LAYERS} conv layers supported")
    desc = L.NetDesc()
    desc.abi_version = L.ABI_VERSION
    desc.n_layers = len(layers)
    desc.head, desc.tail = head, tail
    desc.in_channels = in_channels
    blob = BlobBuilder()
    for i, spec in enumerate(layers):
        d = desc.layers[i]
        w = spec.weight
        if w.dim() != 4 or w.shape[2] != w.shape[3] or int(w.shape[2]) not in (1, 3, 5, 7):
            raise ValueError(f"layer {i + 1}: square kernels of size 1, 3, 5 or 7 are implemented by the engine, got {tuple(w.shape)}")
        if w.shape[2] == 1:                              # 1x1: centre tap of a 3x3 kernel
            w = torch.nn.functional.pad(w, (1, 1, 1, 1))
        d.ksize = int(w.shape[2])
        if w.shape[1] != spec.cin0 + spec.cin1:
            raise ValueError(f"layer {i + 1}: weight has {w.shape[1]} input channels, sources give {spec.cin0 + spec.cin1}")
        d.cin0, d.cin1, d.cout = spec.cin0, spec.cin1, int(w.shape[0])
        d.src0, d.src1, d.skip_src = spec.src0, spec.src1, spec.skip_src
        d.w_off = blob.add(w)
        d.b_off = blob.add(spec.bias) if spec.bias is not None else -1
        for attr, mods in (("pre", spec.pre), ("post", spec.post)):
            mods = [m for m in mods if m.op_name != "identity"]
            if len(mods) > L.MAX_ACTS:
                raise ValueError(f"layer {i + 1}: more than {L.MAX_ACTS} activation slots")
            setattr(d, "n_" + attr, len(mods))
            arr = getattr(d, attr)
            for k, m in enumerate(mods):
                _fill_act(arr[k], m, d.cout, blob)
    return desc, blob.finish()
