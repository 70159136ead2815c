"""B200-native fused inference engine for the FS-UAE image enhancer networks.

Drop-in modules (same names and contracts as the reference's ``model/`` package):
``model_pix_shuffle``, ``model_conv3``, ``model_conv5``, ``activations``; C ABI in
``include/fsuae_enhancer.h`` (``libfsuae_enhancer.so``, built by ``build.build_library``).
"""
from . import activations, model_conv3, model_conv5, model_pix_shuffle  # noqa: F401
from .engine import Engine  # noqa: F401

__all__ = ["activations", "model_pix_shuffle", "model_conv3", "model_conv5", "Engine"]
