"""Thin Python handle over one C-ABI engine (one network, one GPU, one frame size)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib as L


class Engine:
    """Owns a ``fsuae_engine*``.  Device pointers and the CUDA stream come from torch tensors;
    nothing else of torch crosses the boundary."""

    def __init__(self, desc: L.NetDesc, blob: np.ndarray, device: int, precision: int, height: int,
                 width: int, chunk_frames: int = 8):
        self._lib = L.load()
        self._h = C.c_void_p()
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        self.height, self.width, self.device, self.precision = height, width, device, precision
        rc = self._lib.fsuae_engine_create(C.byref(desc), blob.ctypes.data_as(C.POINTER(C.c_float)), blob.size,
                                           device, precision, height, width, chunk_frames, C.byref(self._h))
        if rc != L.OK:
            msg = self._lib.fsuae_last_error(None).decode()
            self._h = C.c_void_p()
            if rc == L.ERR_INVALID:
                raise ValueError(msg)          # reference raises ValueError for bad configs
            raise L.EngineError(rc, msg)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.fsuae_engine_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _check(self, rc: int):
        if rc != L.OK:
            msg = self._lib.fsuae_last_error(self._h).decode()
            if rc == L.ERR_INVALID:
                raise ValueError(msg)
            raise L.EngineError(rc, msg)

    @property
    def variant(self) -> str:
        return self._lib.fsuae_engine_variant(self._h).decode()

    @property
    def device_bytes(self) -> int:
        return int(self._lib.fsuae_engine_device_bytes(self._h))

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.fsuae_engine_last_launch_count(self._h))

    def set_profiling(self, enabled: bool):
        """Bracket every kernel launch with CUDA events (measurement aid; adds a little launch overhead)."""
        self._check(self._lib.fsuae_engine_set_profiling(self._h, int(enabled)))

    def kernel_times(self):
        """[(label, ms), ...] of the last enqueue; synchronise the stream first."""
        out, buf = [], C.create_string_buffer(96)
        i = 0
        while True:
            ms = self._lib.fsuae_engine_kernel_time(self._h, i, buf, 96)
            if ms < 0:
                return out
            out.append((buf.value.decode(), float(ms)))
            i += 1

    def enqueue(self, x: torch.Tensor, out: torch.Tensor, n_frames: int, in_fmt: int, out_fmt: int,
                flags: int = 0, stream: Optional[torch.cuda.Stream] = None):
        """Asynchronous forward of ``n_frames`` frames on ``stream`` (default: torch's current stream)."""
        if not (x.is_cuda and out.is_cuda and x.is_contiguous() and out.is_contiguous()):
            raise ValueError("engine.enqueue needs contiguous CUDA tensors")
        if x.device.index != self.device or out.device.index != self.device:
            raise ValueError("tensor is on a different GPU than the engine")
        st = stream if stream is not None else torch.cuda.current_stream(x.device)
        self._check(self._lib.fsuae_engine_enqueue(self._h, x.data_ptr(), out.data_ptr(), n_frames, in_fmt,
                                                   out_fmt, flags, st.cuda_stream))

    def submit_host(self, x: torch.Tensor, out: torch.Tensor, n_frames: int, in_fmt: int, out_fmt: int,
                    flags: int = 0):
        """Streaming form of :meth:`run_host`: queue upload, forward and download, return at once.  ``x`` and ``out``
        must stay alive (and be pinned for real asynchrony) until :meth:`wait_host` returns."""
        if x.is_cuda or out.is_cuda or not x.is_contiguous() or not out.is_contiguous():
            raise ValueError("engine.submit_host needs contiguous CPU tensors")
        self._check(self._lib.fsuae_engine_submit_host(self._h, x.data_ptr(), out.data_ptr(), n_frames, in_fmt,
                                                       out_fmt, flags))

    def wait_host(self):
        """Block until every frame submitted with :meth:`submit_host` has arrived in its host output buffer."""
        self._check(self._lib.fsuae_engine_wait_host(self._h))

    def run_host(self, x: torch.Tensor, out: torch.Tensor, n_frames: int, in_fmt: int, out_fmt: int,
                 flags: int = 0):
        """Synchronous end-to-end call on HOST tensors (pinned recommended): H2D, forward, D2H."""
        if x.is_cuda or out.is_cuda or not x.is_contiguous() or not out.is_contiguous():
            raise ValueError("engine.run_host needs contiguous CPU tensors")
        self._check(self._lib.fsuae_engine_run_host(self._h, x.data_ptr(), out.data_ptr(), n_frames, in_fmt,
                                                    out_fmt, flags))
