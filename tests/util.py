"""Shared helpers for the test-suite (oracle side = checker only)."""
import os

import numpy as np
import torch

from oracle import enhancer_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_gold(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: z[k] for k in z.files}


def gold_spec(name):
    if name in ("lightweight", "heavyweight"):
        return O.pix_shuffle_preset(name)
    from oracle.gen_golden import KSIZE_SPECS, PROJ_SPECS, VOCAB_SPECS
    for table in (VOCAB_SPECS, PROJ_SPECS, KSIZE_SPECS):
        if name in table:
            return table[name]
    raise KeyError(name)


def trained_pix_shuffle_sd():
    z = np.load(os.path.join(GOLD, "pix_shuffle_trained_fp16.npz"))
    return {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files}


def trained_conv3_sd():
    z = np.load(os.path.join(GOLD, "conv3_trained_fp16.npz"))
    return {k: torch.from_numpy(z[k].astype(np.float32) if z[k].dtype == np.float16 else z[k]) for k in z.files}


def trained_conv3_heavy_sd():
    z = np.load(os.path.join(GOLD, "conv3_heavy_trained_fp16.npz"))
    return {k: torch.from_numpy(z[k].astype(np.float32) if z[k].dtype == np.float16 else z[k]) for k in z.files}


def load_png_rgb(path):
    from PIL import Image
    a = np.asarray(Image.open(path).convert("RGB"))
    return torch.from_numpy(a.copy()).permute(2, 0, 1).unsqueeze(0).contiguous()   # [1,3,H,W] uint8


def load_png_rgba(path):
    from PIL import Image
    a = np.asarray(Image.open(path).convert("RGBA"))
    return torch.from_numpy(a.copy()).unsqueeze(0).contiguous()                     # [1,H,W,4] uint8


def build_pkg_pix_shuffle(spec, sd):
    """Our drop-in Model configured like `spec` and loaded with `sd`."""
    from fs_uae_image_enhancer_project_b200 import model_pix_shuffle
    kw = {f"layer{i + 1}_out_channels": spec.channels[i] for i in range(6)}
    kw.update({f"layer{i + 1}_kernel_size": k for i, k in enumerate(spec.kernel_sizes)})
    for slot, (name, params) in spec.acts.items():
        kw[f"layer{slot[1]}_act{slot[-1]}"] = name
        kw[f"layer{slot[1]}_act{slot[-1]}_params"] = params
    m = model_pix_shuffle.Model(**kw)
    m.load_state_dict(sd, strict=True)
    return m.eval()


def build_pkg_residual_block(name):
    """Our drop-in ResidualFeatureBlock for golden case `name` -> (module, state_dict, acts, golden arrays)."""
    from fs_uae_image_enhancer_project_b200 import residual_feature_block
    from oracle.gen_golden import RESBLOCK_CASES
    ci, cm, co, ks, acts = RESBLOCK_CASES[name]
    g = load_gold(name)
    sd = O.make_residual_block_state_dict(ci, cm, co, ks, acts, int(g["seed"]))
    m = residual_feature_block.ResidualFeatureBlock(ci, cm, co, ks, acts=acts)
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd, acts, g
