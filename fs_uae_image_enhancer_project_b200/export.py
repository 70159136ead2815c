"""Write / read the engine file consumed by ``fsuae_engine_create_from_file`` (C ABI).

This is the deploy-time counterpart of the reference's ONNX export
(``convertion_tools/torch2onnx.py``): instead of an edited ONNX graph for ONNX Runtime, the C side
(FS-UAE) loads one flat file holding the network descriptor and its float32 parameter blob.

Layout: ``b"FSUAEENG" | uint32 abi_version | uint32 blob_floats | fsuae_net_desc | float32[blob_floats]``.
"""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from . import _lib as L
from .descriptor import build_descriptor

MAGIC = b"FSUAEENG"


def export_engine_file(model, path: str) -> int:
    """Serialise a drop-in model (``model_pix_shuffle`` / ``model_conv3`` / ``model_conv5`` instance)."""
    desc, blob = build_descriptor(model._layer_specs(), model._head, model._tail, getattr(model, "_in_channels", 0))
    blob = np.ascontiguousarray(blob, dtype=np.float32)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", L.ABI_VERSION, blob.size))
        f.write(bytes(desc))
        f.write(blob.tobytes())
    return 16 + C.sizeof(desc) + blob.nbytes


def read_engine_file(path: str):
    """-> (NetDesc, float32 blob); raises ValueError on a malformed file."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != MAGIC:
        raise ValueError("not an FSUAE engine file")
    abi, n = struct.unpack_from("<II", data, 8)
    if abi != L.ABI_VERSION:
        raise ValueError(f"engine file ABI {abi}, library ABI {L.ABI_VERSION}")
    desc = L.NetDesc.from_buffer_copy(data[16:16 + C.sizeof(L.NetDesc)])
    blob = np.frombuffer(data, dtype=np.float32, count=n, offset=16 + C.sizeof(L.NetDesc)).copy()
    return desc, blob
