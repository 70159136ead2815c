"""CPU oracle for the FS-UAE image-enhancer forward pass.  TEST INFRASTRUCTURE ONLY.

This module is the *checker* for the CUDA engine.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it; the product package ``fs_uae_image_enhancer_project_b200`` never does.

It restates, in plain functional PyTorch on the CPU (fp32 by default, fp64 on request), the
arithmetic of the reference's hot path.  The arithmetic itself lives in a third-party
dependency of the reference -- PyTorch (the reference pins no version; its shipped ONNX files
say ``producer pytorch 2.5.1``; this container has torch 2.11.0) -- so the restatement is
anchored on the reference's own call sites (all paths relative to /root/reference):

* ``model/model_pix_shuffle.py:227-298``  Model.forward (pix_shuffle)       -> pix_shuffle_forward
* ``model/model_pix_shuffle.py:304-314``  get_model presets                 -> PIX_SHUFFLE_PRESETS
* ``model/model_conv3.py:102-155``        Model.forward (conv3)             -> conv3_forward
* ``model/model_conv5.py:114-151``        Model.forward (conv5)             -> conv5_forward
* ``model/activations.py:6-65, 69-95``    custom activations + registry     -> apply_activation
* ``model/gamma.py:13-15, 31-33``         t**2.2 / t**(1/2.2)               -> gamma_in / gamma_out
* ``model/train.py:57-70``                float inference glue              -> float_pipeline
* ``convertion_tools/torch2onnx.py:224-412, 539-724`` uint8 framebuffer glue -> framebuffer_forward

Parity pinning: ``oracle/gen_golden.py`` imports the *real* reference modules in the build
container and stores input/output vectors under ``tests/golden/``; ``tests/test_oracle.py``
checks this restatement against them, and against the reference's own shipped regression
artefacts (trained ONNX weights + ``model/samples`` -> ``model/*/predicted`` PNGs).
conv5 has no usable shipped artefact -> pinned by live-reference
vectors only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

FRAME_W = 752  # README.md:5
FRAME_H = 576

# --------------------------------------------------------------------------------------
# activation vocabulary (activations.py:69-95)
# --------------------------------------------------------------------------------------

ACTIVATION_NAMES = (
    "identity", "elu", "gelu", "leaky_relu", "mish", "prelu", "relu", "relu6", "sigmoid",
    "silu", "swish", "softplus", "tanh", "log_softmax", "softmax", "scaled_tanh", "telu",
    "sinlu", "biased_relu", "biased_prelu",
)


def _per_channel(p: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """activations.py:44-47 / 61-64: broadcast as [1,C,1,1] iff numel == C, else as-is."""
    if x.dim() == 4 and p.numel() == x.size(1):
        return p.view(1, -1, 1, 1)
    return p


def apply_activation(name: str, x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str,
                     params: Optional[dict] = None) -> torch.Tensor:
    """One activation slot.  ``prefix`` is the module name in the state_dict (e.g. 'l2_act4')."""
    name = name.lower()
    params = params or {}
    if name == "identity":
        return x
    if name == "relu":
        return torch.relu(x)
    if name == "relu6":
        return torch.clamp(x, 0.0, 6.0)
    if name == "tanh":
        return torch.tanh(x)
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name in ("silu", "swish"):
        return x * torch.sigmoid(x)
    if name == "mish":
        return F.mish(x)
    if name == "gelu":
        return F.gelu(x)  # approximate='none'
    if name == "elu":
        return F.elu(x, alpha=float(params.get("alpha", 1.0)))
    if name == "softplus":
        return F.softplus(x, beta=float(params.get("beta", 1.0)),
                          threshold=float(params.get("threshold", 20.0)))
    if name == "leaky_relu":
        return F.leaky_relu(x, negative_slope=float(params.get("negative_slope", 0.01)))
    if name == "prelu":
        return F.prelu(x, sd[prefix + ".weight"].to(x.dtype))
    if name == "softmax":
        return torch.softmax(x, dim=int(params.get("dim", 1)))
    if name == "log_softmax":
        return torch.log_softmax(x, dim=int(params.get("dim", 1)))
    if name == "scaled_tanh":            # activations.py:19-20
        return (torch.tanh(x) + 1.0) * 0.5
    if name == "telu":                   # activations.py:11-12
        return x * torch.tanh(torch.exp(x))
    if name == "sinlu":                  # activations.py:31-32
        a = sd[prefix + ".a"].to(x.dtype)
        b = sd[prefix + ".b"].to(x.dtype)
        return torch.sigmoid(x) * (x + a * torch.sin(b * x))
    if name == "biased_relu":            # activations.py:42-48
        return torch.relu(x - _per_channel(sd[prefix + ".bias"].to(x.dtype), x))
    if name == "biased_prelu":           # activations.py:59-65
        y = x - _per_channel(sd[prefix + ".bias"].to(x.dtype), x)
        return F.prelu(y, sd[prefix + ".prelu.weight"].to(x.dtype))
    raise ValueError(f"Unsupported activation: '{name}'")  # activations.py:123-127


def activation_param_shapes(name: str, params: Optional[dict]) -> Dict[str, Tuple[int, ...]]:
    """state_dict entries (suffix -> shape) an activation slot owns."""
    params = params or {}
    n = int(params.get("num_parameters", 1))
    name = name.lower()
    if name == "sinlu":
        return {".a": (1,), ".b": (1,)}
    if name == "prelu":
        return {".weight": (n,)}
    if name == "biased_relu":
        return {".bias": (n,)}
    if name == "biased_prelu":
        return {".bias": (n,), ".prelu.weight": (n,)}
    return {}


# --------------------------------------------------------------------------------------
# pix_shuffle (model_pix_shuffle.py)
# --------------------------------------------------------------------------------------

@dataclass
class PixShuffleSpec:
    """Constructor arguments of model_pix_shuffle.Model (:20-70) that affect forward()."""
    channels: Tuple[int, int, int, int, int, int] = (36, 36, 36, 36, 36, 36)
    # slot name -> (activation name, params or None); defaults are the ctor defaults (:20-68)
    acts: Dict[str, Tuple[str, Optional[dict]]] = field(default_factory=lambda: {
        "l1_act1": ("identity", None), "l1_act2": ("relu", None),
        "l2_act1": ("mish", None), "l2_act2": ("biased_relu", None),
        "l2_act3": ("tanh", None), "l2_act4": ("relu6", None),
        "l3_act1": ("identity", None), "l3_act2": ("identity", None),
        "l4_act1": ("telu", None), "l4_act2": ("leaky_relu", None),
        "l4_act3": ("tanh", None), "l4_act4": ("identity", None),
        "l5_act1": ("identity", None), "l5_act2": ("identity", None),
        "l6_act1": ("mish", None), "l6_act2": ("prelu", None),
        "l7_act1": ("sinlu", None), "l7_act2": ("prelu", None),
    })
    kernel_sizes: Tuple[int, int, int, int, int, int, int] = (3, 3, 3, 3, 3, 3, 3)    # layer{i}_kernel_size (:21-64), padding (k-1)//2 (:108-115)

    def with_acts(self, **kw) -> "PixShuffleSpec":
        acts = dict(self.acts)
        for k, v in kw.items():
            acts[k] = v if isinstance(v, tuple) else (v, None)
        return PixShuffleSpec(self.channels, acts, self.kernel_sizes)

    def with_kernels(self, *ks) -> "PixShuffleSpec":
        return PixShuffleSpec(self.channels, dict(self.acts), tuple(ks))


def pix_shuffle_preset(name: str) -> PixShuffleSpec:
    """model_pix_shuffle.py:304-314."""
    if name == "lightweight":
        return PixShuffleSpec((36, 36, 72, 72, 36, 36)).with_acts(
            l1_act1="sinlu", l1_act2="relu6",
            l2_act1="telu", l2_act2="identity", l2_act3="sinlu",
            l2_act4=("biased_prelu", {"num_parameters": 36}),
            l4_act1="mish", l4_act2=("biased_prelu", {"num_parameters": 72}),
            l4_act3="tanh", l4_act4="relu",
            l6_act1="mish", l6_act2="relu6",
            l7_act1="identity", l7_act2=("biased_prelu", {"num_parameters": 1}))
    if name == "heavyweight":
        return PixShuffleSpec((36, 36, 108, 108, 36, 36))
    raise ValueError(name)


def pix_shuffle_conv_shapes(spec: PixShuffleSpec) -> List[Tuple[int, int]]:
    """(Cin, Cout) of conv1..conv7 (model_pix_shuffle.py:121-165)."""
    c1, c2, c3, c4, c5, c6 = spec.channels
    return [(12, c1), (c1, c2), (c2, c3), (c3, c4), (c4, c5), (c1 + c5, c6), (c6, 12)]


def pix_shuffle_forward(sd: Dict[str, torch.Tensor], spec: PixShuffleSpec, x: torch.Tensor,
                        dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """model_pix_shuffle.py:227-298, un-fused eval forward.  x: [B,3,H,W] linear RGB."""
    c1, c2, c3, c4, _, _ = spec.channels
    w = lambda k: sd[k].to(dtype)

    def proj(key, t, cin, cout):
        """1x1 projection of a short skip when the channel counts differ (:126-128, :143-145, :250-251, :269-270)."""
        return F.conv2d(t, w(key)) if cin != cout else t

    def act(slot, t):
        name, params = spec.acts[slot]
        return apply_activation(name, t, sd, slot, params)

    def conv(i, t):
        k = w(f"conv{i}.weight")
        return F.conv2d(t, k, w(f"conv{i}.bias"), stride=1, padding=(k.shape[-1] - 1) // 2)        # :108-115

    x = x.to(dtype)
    identity = x                                             # :232
    x = F.pixel_unshuffle(x, 2)                              # :235
    x = act("l1_act2", act("l1_act1", conv(1, x)))           # :238-240
    long_skip = x                                            # :241
    short = x                                                # :244
    x = act("l2_act2", act("l2_act1", conv(2, x)))           # :245-247
    x = proj("skip1_proj_conv.weight", short, c1, c2) + x    # :250-252
    x = act("l2_act4", act("l2_act3", x))                    # :254-255
    x = act("l3_act2", act("l3_act1", conv(3, x)))           # :258-260
    short = x                                                # :263
    x = act("l4_act2", act("l4_act1", conv(4, x)))           # :264-266
    x = proj("skip2_proj_conv.weight", short, c3, c4) + x    # :269-271
    x = act("l4_act4", act("l4_act3", x))                    # :273-274
    x = act("l5_act2", act("l5_act1", conv(5, x)))           # :277-279
    x = torch.cat([long_skip, x], dim=1)                     # :282
    x = act("l6_act2", act("l6_act1", conv(6, x)))           # :283-285
    x = act("l7_act2", act("l7_act1", conv(7, x)))           # :288-290
    x = F.pixel_shuffle(x, 2)                                # :293
    return torch.relu(identity + x)                          # :295-296


# --------------------------------------------------------------------------------------
# residual feature block (residual_feature_block.py:5-55), the bottleneck of model_residual_unet.py
# --------------------------------------------------------------------------------------

def residual_block_forward(sd: Dict[str, torch.Tensor], acts: Dict[str, object], x: torch.Tensor,
                           dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """residual_feature_block.py:44-55: conv1 (1x1) -> conv2 (k x k) -> act1, act2 -> conv3 (1x1) -> act3 -> + identity
    (proj_conv when present) -> act4.  ``acts``: the block's `acts` dict with 'global' / 'channel' already resolved."""
    w = lambda k: sd[k].to(dtype)
    act = lambda key, t: apply_activation(acts[key], t, sd, key, acts.get(key + "_params"))
    x = x.to(dtype)
    identity = x
    t = F.conv2d(x, w("conv1.weight"), w("conv1.bias"))
    k2 = w("conv2.weight")
    t = F.conv2d(t, k2, w("conv2.bias"), padding=(k2.shape[-1] - 1) // 2)
    t = act("act2", act("act1", t))
    t = act("act3", F.conv2d(t, w("conv3.weight"), w("conv3.bias")))
    if "proj_conv.weight" in sd:
        identity = F.conv2d(identity, w("proj_conv.weight"), w("proj_conv.bias"))
    return act("act4", identity + t)


def make_residual_block_state_dict(cin: int, cmid: int, cout: int, ksize: int, acts: Dict[str, object], seed: int):
    rs = np.random.RandomState(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, (ci, co, k) in (("conv1", (cin, cmid, 1)), ("conv2", (cmid, cmid, ksize)), ("conv3", (cmid, cout, 1))):
        b = 1.0 / math.sqrt(ci * k * k)
        sd[f"{name}.weight"] = _u(rs, (co, ci, k, k), -b, b)
        sd[f"{name}.bias"] = _u(rs, (co,), -b, b)
    for key in ("act1", "act2", "act3", "act4"):
        for suffix, shape in activation_param_shapes(acts[key], acts.get(key + "_params")).items():
            lo, hi = (0.6, 1.7) if suffix in (".a", ".b") else ((-0.1, 0.1) if suffix == ".bias" else (0.05, 0.45))
            sd[key + suffix] = _u(rs, shape, lo, hi)
    if cin != cout:
        b = 1.0 / math.sqrt(cin)
        sd["proj_conv.weight"] = _u(rs, (cout, cin, 1, 1), -b, b)
        sd["proj_conv.bias"] = _u(rs, (cout,), -b, b)
    return sd


# --------------------------------------------------------------------------------------
# conv3 / conv5 (deprecated families)
# --------------------------------------------------------------------------------------

CONV3_PRESETS = {"lightweight": (32, 64), "heavyweight": (192, 256)}   # model_conv3.py:206-211
CONV5_PRESETS = {"lightweight": (32, 64), "heavyweight": (64, 128)}    # model_conv5.py:157-162
BN_EPS = 1e-5


def _bn(sd, i, t, dtype):
    return F.batch_norm(t, sd[f"bn{i}.running_mean"].to(dtype), sd[f"bn{i}.running_var"].to(dtype),
                        sd[f"bn{i}.weight"].to(dtype), sd[f"bn{i}.bias"].to(dtype),
                        training=False, eps=BN_EPS)


def conv3_forward(sd: Dict[str, torch.Tensor], x_u8: torch.Tensor,
                  dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """model_conv3.py:102-155.  uint8 [B,4,H,W] planar RGBA -> float [B,4,H,W] (x255, alpha 255)."""
    if x_u8.dtype != torch.uint8 or x_u8.shape[1] != 4:
        raise ValueError("Input tensor must be uint8 with 4 channels (RGBA)")   # :109-110
    t = x_u8[:, :3].float().div(255.0).to(dtype)                                 # :113-121
    t = torch.relu(_bn(sd, 1, F.conv2d(t, sd["conv1.weight"].to(dtype), None, 1, 1), dtype))
    t = torch.relu(_bn(sd, 2, F.conv2d(t, sd["conv2.weight"].to(dtype), None, 1, 1), dtype))
    t = _bn(sd, 3, F.conv2d(t, sd["conv3.weight"].to(dtype), None, 1, 1), dtype)
    t = t.mul(255.0)                                                             # :145
    alpha = torch.full_like(t[:, :1], 255.0)                                     # :149-150
    return torch.cat((t, alpha), dim=1)                                          # :153


def conv5_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor,
                  dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """model_conv5.py:114-151, un-fused eval forward.  float [B,3,H,W] -> same, in (0,1)."""
    cv = lambda i, t: F.conv2d(t, sd[f"conv{i}.weight"].to(dtype), None, 1, 1)
    x = x.to(dtype)
    x = torch.relu(_bn(sd, 1, cv(1, x), dtype))          # :123-125
    skip = x
    x = torch.relu(skip + _bn(sd, 2, cv(2, x), dtype))   # :128-132
    x = torch.relu(_bn(sd, 3, cv(3, x), dtype))          # :135-137
    skip = x
    x = torch.relu(skip + _bn(sd, 4, cv(4, x), dtype))   # :140-144
    return torch.sigmoid(_bn(sd, 5, cv(5, x), dtype))    # :147-149


# --------------------------------------------------------------------------------------
# gamma + framebuffer glue
# --------------------------------------------------------------------------------------

def gamma_in(t: torch.Tensor) -> torch.Tensor:
    """gamma.py:13-15 srgb_to_linear_approx."""
    return t ** 2.2


def gamma_out(t: torch.Tensor) -> torch.Tensor:
    """gamma.py:31-33 linear_to_srgb_approx."""
    return t ** (1.0 / 2.2)


def to_u8_trunc(t01: torch.Tensor) -> torch.Tensor:
    """clamp -> x255 -> truncating cast (train.py:70-73 via ToPILImage; torch2onnx.py:562-632)."""
    return (t01.clamp(0.0, 1.0) * 255.0).to(torch.uint8)


def float_pipeline(sd, spec: PixShuffleSpec, rgb_u8_nchw: torch.Tensor,
                   dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """train.py:57-73 inference_on_directory: uint8 sRGB [B,3,H,W] -> uint8 sRGB [B,3,H,W]."""
    x = gamma_in(rgb_u8_nchw.to(torch.float32) / 255.0)
    y = pix_shuffle_forward(sd, spec, x, dtype).to(torch.float32)
    return to_u8_trunc(gamma_out(y))


def framebuffer_forward(sd, spec: PixShuffleSpec, rgba_u8_nhwc: torch.Tensor, crop16: bool = False,
                        dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Deployed contract (torch2onnx.py:184-768), evaluated in fp32 instead of the graph's fp16:
    uint8 [B,H,W,4] RGBA -> uint8 [B,H,W,4], alpha 255; ``crop16`` = newer exporter
    (:299-355 crop x in [16,W), :634-674 pad 16 black columns on the left)."""
    rgb = rgba_u8_nhwc[..., :3].permute(0, 3, 1, 2)              # :225-297
    if crop16:
        rgb = rgb[..., 16:]                                      # :299-355
    out = float_pipeline(sd, spec, rgb.contiguous(), dtype)      # :358-412, model, :539-632
    if crop16:
        out = F.pad(out, (16, 0, 0, 0), value=0)                 # :634-674
    B, _, H, W = out.shape
    alpha = torch.full((B, 1, H, W), 255, dtype=torch.uint8)     # :677-724
    return torch.cat([out, alpha], dim=1).permute(0, 2, 3, 1).contiguous()


# --------------------------------------------------------------------------------------
# deterministic weights (numpy RandomState: stream is frozen by numpy policy)
# --------------------------------------------------------------------------------------

def _u(rs: np.random.RandomState, shape, lo, hi) -> torch.Tensor:
    return torch.from_numpy(rs.uniform(lo, hi, size=shape).astype(np.float32))


def make_pix_shuffle_state_dict(spec: PixShuffleSpec, seed: int) -> Dict[str, torch.Tensor]:
    """Seeded weights with the key set / shapes of the reference state_dict."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, torch.Tensor] = {}
    for i, (ci, co) in enumerate(pix_shuffle_conv_shapes(spec), start=1):
        ks = spec.kernel_sizes[i - 1]
        k = 1.0 / math.sqrt(ci * ks * ks)
        sd[f"conv{i}.weight"] = _u(rs, (co, ci, ks, ks), -k, k)
        sd[f"conv{i}.bias"] = _u(rs, (co,), -k, k)
    for slot in sorted(spec.acts):
        name, params = spec.acts[slot]
        for suffix, shape in activation_param_shapes(name, params).items():
            if suffix in (".a", ".b"):
                sd[slot + suffix] = _u(rs, shape, 0.6, 1.7)
            elif suffix == ".bias":
                sd[slot + suffix] = _u(rs, shape, -0.1, 0.1)
            else:
                sd[slot + suffix] = _u(rs, shape, 0.05, 0.45)
    c1, c2, c3, c4, _, _ = spec.channels          # drawn last: the streams of the projection-free specs stay as they were
    if c1 != c2:
        sd["skip1_proj_conv.weight"] = _u(rs, (c2, c1, 1, 1), -1.0 / math.sqrt(c1), 1.0 / math.sqrt(c1))
    if c3 != c4:
        sd["skip2_proj_conv.weight"] = _u(rs, (c4, c3, 1, 1), -1.0 / math.sqrt(c3), 1.0 / math.sqrt(c3))
    return sd


def make_bn_state_dict(channels: Sequence[Tuple[int, int]], seed: int) -> Dict[str, torch.Tensor]:
    """conv{i}.weight + randomised BN statistics (random-init BN hides folding bugs)."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, torch.Tensor] = {}
    for i, (ci, co) in enumerate(channels, start=1):
        k = 1.0 / math.sqrt(ci * 9)
        sd[f"conv{i}.weight"] = _u(rs, (co, ci, 3, 3), -k, k)
        sd[f"bn{i}.weight"] = _u(rs, (co,), 0.6, 1.4)
        sd[f"bn{i}.bias"] = _u(rs, (co,), -0.2, 0.2)
        sd[f"bn{i}.running_mean"] = _u(rs, (co,), -0.2, 0.2)
        sd[f"bn{i}.running_var"] = _u(rs, (co,), 0.5, 1.5)
        sd[f"bn{i}.num_batches_tracked"] = torch.tensor(7, dtype=torch.long)
    return sd


def conv3_channels(preset: str):
    a, b = CONV3_PRESETS[preset]
    return [(3, a), (a, b), (b, 3)]


def conv5_channels(preset: str):
    a, b = CONV5_PRESETS[preset]
    return [(3, a), (a, a), (a, b), (b, b), (b, 3)]


# --------------------------------------------------------------------------------------
# synthetic Amiga framebuffers (README.md:7-10; util.py:335-348;
# rgb444_flat_image_generator.py:28-30)
# --------------------------------------------------------------------------------------

PIXEL_MODES = {"lores": (2, 2), "lores_laced": (1, 2), "hires": (2, 1), "hires_laced": (1, 1)}
# name -> (rows replicated sy, columns replicated sx)


def synth_framebuffers(n: int, seed: int, h: int = FRAME_H, w: int = FRAME_W,
                       modes: Sequence[str] = ("lores", "lores_laced", "hires", "hires_laced"),
                       scale: int = 17) -> torch.Tensor:
    """n RGBA uint8 frames [n,h,w,4]; frame i uses pixel mode ``modes[i * len(modes) // n]``
    (equal contiguous groups); RGB444 values q*scale (17 = Amiga expansion), alpha 255."""
    g = torch.Generator().manual_seed(seed)
    out = torch.empty((n, h, w, 4), dtype=torch.uint8)
    out[..., 3] = 255
    for i in range(n):
        sy, sx = PIXEL_MODES[modes[(i * len(modes)) // n]]
        q = torch.randint(0, 16, (3, (h + sy - 1) // sy, (w + sx - 1) // sx), generator=g,
                          dtype=torch.int64)
        q = q.repeat_interleave(sy, dim=1).repeat_interleave(sx, dim=2)[:, :h, :w]
        out[i, :, :, :3] = (q * scale).to(torch.uint8).permute(1, 2, 0)
    return out


def psnr(a: torch.Tensor, b: torch.Tensor, peak: float) -> float:
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)


# --------------------------------------------------------------------------------------
# input side: colour-depth grid quantisation, pixel-mode replication, counter-hashed synthetic frames
# (dataset_generator/quantize.py:464-473, 512-521; dataset_generator/util.py:318-350)
# --------------------------------------------------------------------------------------

def quantize_grid(img: np.ndarray, color_space: str) -> np.ndarray:
    """quantize.py:512-521 with dithering_method='none', target_palette_size=None: floor onto the colour grid.
    img: uint8 [..., 3]."""
    t = img.astype(np.float64)
    if color_space == "RGB444":
        t = np.floor(t / 16) * 16
    elif color_space == "RGB666":
        t = np.floor(t / 4) * 4
    elif color_space == "RGB565":
        t = t.copy()
        t[..., 0] = np.floor(t[..., 0] / 8) * 8
        t[..., 1] = np.floor(t[..., 1] / 4) * 4
        t[..., 2] = np.floor(t[..., 2] / 8) * 8
    elif color_space == "RGB555":
        t = np.floor(t / 8) * 8
    elif color_space != "RGB888":
        raise ValueError(f"Invalid color_space '{color_space}'")
    return np.clip(t, 0, 255).astype(np.uint8)


def post_resolution_style(img: np.ndarray, style: str) -> np.ndarray:
    """util.py:318-350: nearest-neighbour replication back to display resolution.  img: [h, w, c]."""
    sy, sx = PIXEL_MODES[style]
    return np.repeat(np.repeat(img, sy, axis=0), sx, axis=1)


def quantize_frames(img: np.ndarray, color_space: str, style: str, expand17: bool = False) -> np.ndarray:
    """[B,h,w,3|4] uint8 -> RGBA [B,h*sy,w*sx,4]: what fsuae_quantize_frames computes."""
    q = quantize_grid(img[..., :3], color_space)
    if expand17 and color_space == "RGB444":
        q = (q >> 4) * 17                                   # rgb444_flat_image_generator.py:28-30
    out = np.stack([post_resolution_style(f, style) for f in q]) if len(q) else \
        np.zeros((0, img.shape[1] * PIXEL_MODES[style][0], img.shape[2] * PIXEL_MODES[style][1], 3), np.uint8)
    alpha = np.full(out.shape[:3] + (1,), 255, np.uint8)
    return np.concatenate([out.astype(np.uint8), alpha], axis=3)


BAYER_MATRICES = {      # quantize.py:336-357
    "bayer2x2": np.array([[0, 2], [3, 1]], dtype=np.int32),
    "bayer4x4": np.array([[0, 8, 2, 10], [12, 4, 14, 6], [3, 11, 1, 9], [15, 7, 13, 5]], dtype=np.int32),
    "bayer8x8": np.array([[0, 32, 8, 40, 2, 34, 10, 42], [48, 16, 56, 24, 50, 18, 58, 26], [12, 44, 4, 36, 14, 46, 6, 38],
                          [60, 28, 52, 20, 62, 30, 54, 22], [3, 35, 11, 43, 1, 33, 9, 41], [51, 19, 59, 27, 49, 17, 57, 25],
                          [15, 47, 7, 39, 13, 45, 5, 37], [63, 31, 55, 23, 61, 29, 53, 21]], dtype=np.int32),
}


def dither_palette(img: np.ndarray, palette: np.ndarray, method: str) -> np.ndarray:
    """quantize.py:137-331 + :529-537 restated in vectorised numpy (float64 like the reference): img uint8 [h,w,3],
    palette uint8 [N,3] -> uint8 [h,w,3].  'none' = nearest palette colour (np.argmin: first minimum), 'checkerboard'
    (:137-229), 'bayer2x2|4x4|8x8' (:232-331)."""
    h, w, _ = img.shape
    n = palette.shape[0]
    if n == 0:
        return np.zeros_like(img)
    if n == 1:
        return np.broadcast_to(palette[0], img.shape).copy()
    px = img.astype(np.float64).reshape(-1, 1, 3)
    pal = palette.astype(np.float64)
    d = ((px[:, :, 0] - pal[None, :, 0]) ** 2 + (px[:, :, 1] - pal[None, :, 1]) ** 2) + (px[:, :, 2] - pal[None, :, 2]) ** 2
    i1 = np.argmin(d, axis=1)                                   # first index of the minimum, as the strict '<' loop
    d1 = d[np.arange(len(d)), i1]
    if method == "none":
        return palette[i1].reshape(img.shape)
    d_rest = d.copy()
    d_rest[np.arange(len(d)), i1] = np.inf
    i2 = np.argmin(d_rest, axis=1)
    yy, xx = np.divmod(np.arange(h * w), w)
    if method == "checkerboard":
        chosen = np.where(d1 == 0.0, i1, np.where((xx + yy) % 2 == 0, i1, i2))
        return palette[chosen].reshape(img.shape)
    mat = BAYER_MATRICES[method]
    msz = mat.shape[0]
    thr = (mat.astype(np.float64) / (msz * msz))[yy % msz, xx % msz]
    lum = lambda c: (c[:, 0] * 0.2126 + c[:, 1] * 0.7152) + c[:, 2] * 0.0722
    lp, l1, l2 = lum(px[:, 0, :]), lum(pal[i1]), lum(pal[i2])
    swap = l1 > l2
    dark, light = np.where(swap, i2, i1), np.where(swap, i1, i2)
    la, lb = np.where(swap, l2, l1), np.where(swap, l1, l2)
    with np.errstate(divide="ignore", invalid="ignore"):
        frac = np.where(np.abs(lb - la) < 1e-6, 0.0, (lp - la) / (lb - la))
    frac = np.maximum(0.0, np.minimum(1.0, frac))
    chosen = np.where(d1 == 0.0, i1, np.where(frac > thr, light, dark))
    return palette[chosen].reshape(img.shape)


def _mix64(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


SYNTH_STYLE_OF_FRAME = ("lores", "lores_laced", "hires", "hires_laced")


def synth_rgb444_frames(n: int, h: int, w: int, seed: int, first_frame: int = 0, expand17: bool = True) -> np.ndarray:
    """The device-side synthetic stream (fsuae_synth_rgb444_frames): frame g = first_frame + i uses pixel mode g & 3,
    every sy x sx cell one counter-hashed 12-bit colour."""
    out = np.empty((n, h, w, 4), np.uint8)
    out[..., 3] = 255
    with np.errstate(over="ignore"):
        for i in range(n):
            g = np.uint64(first_frame + i)
            sy, sx = PIXEL_MODES[SYNTH_STYLE_OF_FRAME[int(g) & 3]]
            cy, cx = np.meshgrid(np.arange(h, dtype=np.uint64) // np.uint64(sy), np.arange(w, dtype=np.uint64) // np.uint64(sx),
                                 indexing="ij")
            key = _mix64(np.array(np.uint64(seed & (2 ** 64 - 1)) + np.uint64(0x9E3779B97F4A7C15) * (g + np.uint64(1))))
            r = _mix64(key ^ (cy * np.uint64(65536) + cx))
            for c in range(3):
                q = ((r >> np.uint64(20 * c + 4)) & np.uint64(15)).astype(np.uint8)
                out[i, :, :, c] = q * (17 if expand17 else 16)
    return out
