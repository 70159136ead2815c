"""Activation vocabulary of the enhancer networks, as *parameter holders* for the fused engine.

Mirrors the interface of the reference's ``model/activations.py`` (``get_activation(name, params)``
at :99-171 over the registry at :69-95; custom modules TeLU :6-12, ScaledTanh :14-20, SinLU :22-32,
BiasedReLU :34-48, BiasedPReLU :50-65): same names, same constructor arguments, same parameter
names/shapes/initialisation, hence the same ``state_dict`` keys.  The arithmetic itself is NOT
here: every slot is executed inside the CUDA epilogue of the convolution it follows
(csrc/common.cuh ``act_apply_accurate`` and the bf16 epilogues), selected by the op-code each
module reports through ``engine_op``.  Calling one of these modules on its own therefore raises.
"""
from __future__ import annotations

import inspect
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from ._lib import ACT


class FusedActivation(nn.Module):
    """Base: an activation slot that only exists inside the fused CUDA pass."""
    op_name = "identity"

    def engine_op(self, channels: int) -> Tuple[str, List[torch.Tensor]]:
        """(op name, up to two parameter tensors each of numel 1 or ``channels``)."""
        return self.op_name, []

    def forward(self, x):  # noqa: D401
        raise RuntimeError(
            f"{type(self).__name__} is fused into the B200 engine's conv epilogue and has no "
            "standalone (CPU/PyTorch) implementation; call the enclosing model instead.")


def _plain(op: str, doc: str):
    return type(op.title().replace("_", ""), (FusedActivation,), {"op_name": op, "__doc__": doc})


Identity = _plain("identity", "x")
ReLU = _plain("relu", "max(x, 0)")
ReLU6 = _plain("relu6", "clamp(x, 0, 6)")
Tanh = _plain("tanh", "tanh(x)")
Sigmoid = _plain("sigmoid", "1 / (1 + exp(-x))")
SiLU = _plain("silu", "x * sigmoid(x)")
Mish = _plain("mish", "x * tanh(softplus(x))")
GELU = _plain("gelu", "0.5 x (1 + erf(x / sqrt 2))")
ScaledTanh = _plain("scaled_tanh", "(tanh(x) + 1) / 2   (reference activations.py:14-20)")
TeLU = _plain("telu", "x * tanh(exp(x))   (reference activations.py:6-12)")


class ELU(FusedActivation):
    op_name = "elu"

    def __init__(self, alpha: float = 1.0, inplace: bool = False):
        super().__init__()
        self.alpha = float(alpha)

    def engine_op(self, channels):
        return "elu", [torch.tensor([self.alpha])]


class LeakyReLU(FusedActivation):
    op_name = "leaky_relu"

    def __init__(self, negative_slope: float = 0.01, inplace: bool = False):
        super().__init__()
        self.negative_slope = float(negative_slope)

    def engine_op(self, channels):
        return "leaky_relu", [torch.tensor([self.negative_slope])]


class Softplus(FusedActivation):
    op_name = "softplus"

    def __init__(self, beta: float = 1.0, threshold: float = 20.0):
        super().__init__()
        self.beta, self.threshold = float(beta), float(threshold)

    def engine_op(self, channels):
        return "softplus", [torch.tensor([self.beta]), torch.tensor([self.threshold])]


class _ChannelSoftmax(FusedActivation):
    def __init__(self, dim: Optional[int] = None):
        super().__init__()
        self.dim = dim

    def engine_op(self, channels):
        if self.dim != 1:
            raise ValueError(f"{self.op_name}: the engine reduces over channels only (dim=1), got dim={self.dim}")
        return self.op_name, []


class Softmax(_ChannelSoftmax):
    op_name = "softmax"


class LogSoftmax(_ChannelSoftmax):
    op_name = "log_softmax"


class PReLU(FusedActivation):
    """Same parameter as ``torch.nn.PReLU``: ``weight[num_parameters]`` initialised to ``init``."""
    op_name = "prelu"

    def __init__(self, num_parameters: int = 1, init: float = 0.25):
        super().__init__()
        self.num_parameters = num_parameters
        self.weight = nn.Parameter(torch.full((num_parameters,), float(init)))

    def engine_op(self, channels):
        return "prelu", [self.weight]


class SinLU(FusedActivation):
    """sigmoid(x) * (x + a sin(b x)), learnable scalars a, b = 1 (reference activations.py:22-32)."""
    op_name = "sinlu"

    def __init__(self):
        super().__init__()
        self.a = nn.Parameter(torch.ones(1))
        self.b = nn.Parameter(torch.ones(1))

    def engine_op(self, channels):
        return "sinlu", [self.a, self.b]


def _channel_bias(num_parameters: int) -> nn.Parameter:
    p = nn.Parameter(torch.empty(num_parameters))
    nn.init.uniform_(p, a=-0.1, b=0.1)
    return p


class BiasedReLU(FusedActivation):
    """relu(x - bias_c) (reference activations.py:34-48)."""
    op_name = "biased_relu"

    def __init__(self, num_parameters: int = 1):
        super().__init__()
        self.bias = _channel_bias(num_parameters)

    def engine_op(self, channels):
        return "biased_relu", [self.bias]


class BiasedPReLU(FusedActivation):
    """prelu(x - bias_c) with the slope held by a nested ``prelu`` module
    (reference activations.py:50-65) so the keys are ``bias`` and ``prelu.weight``."""
    op_name = "biased_prelu"

    def __init__(self, num_parameters: int = 1, init: float = 0.25):
        super().__init__()
        self.bias = _channel_bias(num_parameters)
        self.prelu = PReLU(num_parameters=num_parameters, init=init)

    def engine_op(self, channels):
        return "biased_prelu", [self.bias, self.prelu.weight]


ACTIVATION_REGISTRY: Dict[str, type] = {
    "identity": Identity, "elu": ELU, "gelu": GELU, "leaky_relu": LeakyReLU, "mish": Mish, "prelu": PReLU,
    "relu": ReLU, "relu6": ReLU6, "sigmoid": Sigmoid, "silu": SiLU, "swish": SiLU, "softplus": Softplus,
    "tanh": Tanh, "log_softmax": LogSoftmax, "softmax": Softmax, "scaled_tanh": ScaledTanh, "telu": TeLU,
    "sinlu": SinLU, "biased_relu": BiasedReLU, "biased_prelu": BiasedPReLU,
}
assert all(cls.op_name in ACT for cls in ACTIVATION_REGISTRY.values())


def get_activation(activation_name: str, params: Optional[dict] = None, inplace: bool = False) -> FusedActivation:
    """Create the activation slot ``activation_name`` (case-insensitive).

    Error behaviour follows the reference factory: unknown name -> ``ValueError`` listing the
    supported names; constructor arguments the activation does not take -> ``TypeError``.
    Softmax / LogSoftmax default to ``dim=1`` when no params are given.  ``inplace`` is accepted
    and meaningless (everything is in registers inside the fused pass).
    """
    key = activation_name.lower()
    if key not in ACTIVATION_REGISTRY:
        raise ValueError(f"Unsupported activation: '{activation_name}'. "
                         f"Supported activations are: {list(ACTIVATION_REGISTRY.keys())}")
    cls = ACTIVATION_REGISTRY[key]
    kwargs = dict(params) if params is not None else {}
    if params is None and issubclass(cls, _ChannelSoftmax):
        kwargs["dim"] = 1
    accepted = inspect.signature(cls.__init__).parameters
    if inplace and "inplace" in accepted:
        kwargs.setdefault("inplace", True)
    try:
        return cls(**kwargs)
    except TypeError as exc:
        raise TypeError(f"Error instantiating activation '{activation_name}' with parameters {kwargs}. "
                        f"Check if the parameters match the constructor arguments of {cls.__name__}. "
                        f"Original error: {exc}")
