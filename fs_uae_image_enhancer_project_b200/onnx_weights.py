"""Pull trained weights out of the reference's shipped ``.onnx`` graphs.

The reference ships no ``.pt``/``.pth``; trained weights survive only as fp16 initialisers inside
``model/model_*/**.onnx`` (exported by ``convertion_tools/torch2onnx.py:78-123``).  The ``onnx``
package is not a dependency here, so this is a ~60-line reader of the protobuf wire format
(field numbers from onnx.proto3: ModelProto.graph=7; GraphProto.node=1/.initializer=5;
TensorProto.dims=1, .data_type=2, .name=8, .raw_data=9; NodeProto.input=1/.output=2/.op_type=4).
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Tuple

import numpy as np
import torch

_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64, 10: np.float16,
           11: np.float64}


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """Yield (field_number, wire_type, value) for one message."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            val, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, val


def _tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype, name, raw = 1, "", b""
    for fno, wt, val in _fields(buf):
        if fno == 1:
            if wt == 0:
                dims.append(val)
            else:                      # packed repeated int64
                p = 0
                while p < len(val):
                    d, p = _varint(val, p)
                    dims.append(d)
        elif fno == 2:
            dtype = val
        elif fno == 8:
            name = val.decode()
        elif fno == 9:
            raw = val
    arr = np.frombuffer(raw, dtype=_DTYPES[dtype]).reshape(dims) if raw else np.zeros(dims, _DTYPES[dtype])
    return name, arr


def read_onnx_initializers(path: str) -> Tuple[Dict[str, np.ndarray], List[Tuple[str, List[str], List[str]]]]:
    """Return ({initializer name: array}, [(op_type, inputs, outputs), ...]) of an ONNX file."""
    with open(path, "rb") as f:
        model = f.read()
    graph = next(val for fno, wt, val in _fields(model) if fno == 7 and wt == 2)
    inits: Dict[str, np.ndarray] = {}
    nodes: List[Tuple[str, List[str], List[str]]] = []
    for fno, wt, val in _fields(graph):
        if fno == 5 and wt == 2:
            name, arr = _tensor(val)
            inits[name] = arr
        elif fno == 1 and wt == 2:
            ins, outs, op = [], [], ""
            for f2, w2, v2 in _fields(val):
                if f2 == 1:
                    ins.append(v2.decode())
                elif f2 == 2:
                    outs.append(v2.decode())
                elif f2 == 4:
                    op = v2.decode()
            nodes.append((op, ins, outs))
    return inits, nodes


def pix_shuffle_state_dict_from_onnx(path: str) -> Dict[str, torch.Tensor]:
    """state_dict for ``get_model('lightweight')`` from ``model/model_pix_shuffle/pix_shuffle.onnx``.

    Exporter quirks handled: PReLU slopes are anonymous ``onnx::PRelu_*`` initialisers of shape
    [C,1,1] consumed in graph order by the three BiasedPReLU slots (l2_act4, l4_act2, l7_act2);
    per-channel biases that were constant-folded to a single value are stored as [1] and expanded.
    """
    inits, nodes = read_onnx_initializers(path)
    sd: Dict[str, torch.Tensor] = {}
    for name, arr in inits.items():
        if name.startswith(("conv", "l")) and "." in name and not name.startswith("onnx::"):
            sd[name] = torch.from_numpy(arr.astype(np.float32))
    prelu_slopes = [inits[ins[1]] for op, ins, _ in nodes if op == "PRelu" and ins[1] in inits]
    slots = [("l2_act4", 36), ("l4_act2", 72), ("l7_act2", 1)]
    if len(prelu_slopes) != len(slots):
        raise ValueError(f"expected {len(slots)} PRelu nodes, found {len(prelu_slopes)}")
    for (slot, n), slope in zip(slots, prelu_slopes):
        s = torch.from_numpy(slope.astype(np.float32)).reshape(-1)
        sd[f"{slot}.prelu.weight"] = s.expand(n).clone() if s.numel() == 1 else s
        b = sd.get(f"{slot}.bias")
        if b is None:
            raise ValueError(f"missing {slot}.bias in {path}")
        sd[f"{slot}.bias"] = b.reshape(-1).expand(n).clone() if b.numel() == 1 else b.reshape(-1)
    return sd


def conv3_state_dict_from_onnx(path: str) -> Dict[str, torch.Tensor]:
    """state_dict for ``model_conv3`` from ``conv3.onnx`` / ``conv3_heavy.onnx``.  The exporter had
    BN folded into the convs (``conv{1,2}.0.{weight,bias}``, ``conv3.{weight,bias}``), so BN is
    rebuilt as the identity-with-bias: gamma=1, beta=b, mean=0, var=1-eps."""
    inits, _ = read_onnx_initializers(path)
    sd: Dict[str, torch.Tensor] = {}
    for i, stem in ((1, "conv1.0"), (2, "conv2.0"), (3, "conv3")):
        w = torch.from_numpy(inits[f"{stem}.weight"].astype(np.float32))
        b = torch.from_numpy(inits[f"{stem}.bias"].astype(np.float32))
        co = w.shape[0]
        sd[f"conv{i}.weight"] = w
        sd[f"bn{i}.weight"] = torch.ones(co)
        sd[f"bn{i}.bias"] = b
        sd[f"bn{i}.running_mean"] = torch.zeros(co)
        sd[f"bn{i}.running_var"] = torch.full((co,), 1.0 - 1e-5)
        sd[f"bn{i}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return sd
