"""ctypes binding of the C ABI in include/fsuae_enhancer.h.

The shared library is built in-tree by ``fs_uae_image_enhancer_project_b200.build``.  There is no
CPU or PyTorch fallback: if the library is missing or no sm_100 GPU is present, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

ABI_VERSION = 2
MAX_LAYERS = 16
MAX_ACTS = 4

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE = range(5)

ACT = {name: i for i, name in enumerate([
    "identity", "relu", "relu6", "tanh", "sigmoid", "silu", "mish", "gelu", "elu", "softplus",
    "leaky_relu", "prelu", "scaled_tanh", "telu", "sinlu", "biased_relu", "biased_prelu", "softmax",
    "log_softmax"])}
ACT["swish"] = ACT["silu"]

HEAD_PLAIN, HEAD_UNSHUFFLE2, HEAD_FEATURES = 0, 1, 2
TAIL_PLAIN, TAIL_SHUFFLE2_RESIDUAL_RELU, TAIL_SCALE255_ALPHA, TAIL_FEATURES = 0, 1, 2, 3
PREC_FP32, PREC_BF16, PREC_FP16 = 0, 1, 2
MAX_CHUNK_FRAMES = 1024
FMT_F32_NCHW3, FMT_U8_NHWC4, FMT_U8_NCHW4, FMT_F32_NCHW4, FMT_F32_NCHW = 0, 1, 2, 3, 4
FLAG_GAMMA_IN, FLAG_GAMMA_OUT, FLAG_CROP16 = 1, 2, 4


class ActDesc(C.Structure):
    _fields_ = [("op", C.c_int32), ("n0", C.c_int32), ("p0_off", C.c_int32),
                ("n1", C.c_int32), ("p1_off", C.c_int32)]


class LayerDesc(C.Structure):
    _fields_ = [("cin0", C.c_int32), ("cin1", C.c_int32), ("cout", C.c_int32),
                ("src0", C.c_int32), ("src1", C.c_int32), ("skip_src", C.c_int32),
                ("w_off", C.c_int32), ("b_off", C.c_int32),
                ("n_pre", C.c_int32), ("n_post", C.c_int32),
                ("pre", ActDesc * MAX_ACTS), ("post", ActDesc * MAX_ACTS),
                ("ksize", C.c_int32), ("reserved", C.c_int32)]


class NetDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_layers", C.c_int32),
                ("head", C.c_int32), ("tail", C.c_int32), ("in_channels", C.c_int32), ("reserved", C.c_int32),
                ("layers", LayerDesc * MAX_LAYERS)]


EXPORTS = {
    # name: (restype, argtypes)
    "fsuae_abi_version": (C.c_int, []),
    "fsuae_engine_create": (C.c_int, [C.POINTER(NetDesc), C.POINTER(C.c_float), C.c_size_t, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "fsuae_engine_create_from_file": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.POINTER(C.c_void_p)]),
    "fsuae_engine_destroy": (C.c_int, [C.c_void_p]),
    "fsuae_engine_enqueue": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_uint32, C.c_void_p]),
    "fsuae_engine_run_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_uint32]),
    "fsuae_engine_submit_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                           C.c_uint32]),
    "fsuae_engine_wait_host": (C.c_int, [C.c_void_p]),
    "fsuae_quantize_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p]),
    "fsuae_dither_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                      C.c_void_p]),
    "fsuae_synth_rgb444_frames": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_int, C.c_void_p]),
    "fsuae_engine_device_bytes": (C.c_size_t, [C.c_void_p]),
    "fsuae_engine_last_launch_count": (C.c_int64, [C.c_void_p]),
    "fsuae_engine_variant": (C.c_char_p, [C.c_void_p]),
    "fsuae_engine_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "fsuae_engine_kernel_time": (C.c_float, [C.c_void_p, C.c_int, C.c_char_p, C.c_int]),
    "fsuae_last_error": (C.c_char_p, [C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libfsuae_enhancer.so and bind every symbol the header declares."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m fs_uae_image_enhancer_project_b200.build` "
            "(nvcc, sm_100a).  This engine has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.fsuae_abi_version() != ABI_VERSION:
        raise RuntimeError("libfsuae_enhancer.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fsuae engine error {code}: {msg}")
        self.code = code
