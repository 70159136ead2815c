#!/usr/bin/env python3
"""Generate tests/golden/* by running the REAL reference modules.  TEST INFRASTRUCTURE ONLY.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box);
the vectors it writes are committed.  Recipe: put
/root/reference/model on sys.path and pre-seed ``loss_vgg`` / ``loss_ssim`` with stub modules
(the real ones need kornia, a VGG16 download and a file that is not in the repo).

    python oracle/gen_golden.py            # rewrites tests/golden/
"""
from __future__ import annotations

import os
import shutil
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import enhancer_oracle as O  # noqa: E402
from fs_uae_image_enhancer_project_b200 import onnx_weights  # noqa: E402


def import_reference():
    sys.path.insert(0, os.path.join(REF, "model"))
    for mod, cls in (("loss_vgg", "PerceptualLoss"), ("loss_ssim", "SSIMLoss")):
        m = types.ModuleType(mod)

        class _Stub(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()
        setattr(m, cls, _Stub)
        sys.modules[mod] = m
    import model_pix_shuffle, model_conv3, model_conv5, gamma  # noqa
    return model_pix_shuffle, model_conv3, model_conv5, gamma


def ref_pix_shuffle(mps, spec: O.PixShuffleSpec):
    c = spec.channels
    kw = {f"layer{i + 1}_out_channels": c[i] for i in range(6)}
    kw.update({f"layer{i + 1}_kernel_size": k for i, k in enumerate(spec.kernel_sizes)})
    for slot, (name, params) in spec.acts.items():
        layer, idx = slot[1], slot[-1]
        kw[f"layer{layer}_act{idx}"] = name
        kw[f"layer{layer}_act{idx}_params"] = params
    return mps.Model(**kw).eval()


# activation-vocabulary coverage: every registry name appears in some slot
VOCAB_SPECS = {
    "vocab_a": O.PixShuffleSpec((24, 24, 40, 40, 16, 24)).with_acts(
        l1_act1="gelu", l1_act2=("leaky_relu", {"negative_slope": 0.05}),
        l2_act1="silu", l2_act2=("biased_relu", {"num_parameters": 24}),
        l2_act3="sigmoid", l2_act4=("prelu", {"num_parameters": 24}),
        l3_act1="elu", l3_act2="relu",
        l4_act1="softplus", l4_act2=("biased_prelu", {"num_parameters": 1}),
        l4_act3="scaled_tanh", l4_act4="relu6",
        l5_act1="swish", l5_act2="identity",
        l6_act1="telu", l6_act2=("biased_relu", {"num_parameters": 1}),
        l7_act1="tanh", l7_act2="identity"),
    "vocab_b": O.PixShuffleSpec((16, 16, 32, 32, 16, 16)).with_acts(
        l1_act1="mish", l1_act2="identity",
        l2_act1="sinlu", l2_act2="relu",
        l2_act3="softmax", l2_act4="identity",
        l3_act1="log_softmax", l3_act2="tanh",
        l4_act1=("elu", {"alpha": 0.7}), l4_act2=("prelu", {"num_parameters": 1}),
        l4_act3="gelu", l4_act4=("leaky_relu", None),
        l6_act1="sigmoid", l6_act2="relu6",
        l7_act1="sinlu", l7_act2=("prelu", None)),
}


# channel plans whose short skips need the 1x1 projections (model_pix_shuffle.py:126-128, 143-145); arbitrary widths
PROJ_SPECS = {
    "proj_a": O.PixShuffleSpec((24, 40, 40, 56, 20, 28)).with_acts(
        l1_act1="sinlu", l1_act2="relu6",
        l2_act1="telu", l2_act2="identity", l2_act3="sinlu", l2_act4=("biased_prelu", {"num_parameters": 40}),
        l4_act1="mish", l4_act2=("biased_prelu", {"num_parameters": 56}), l4_act3="tanh", l4_act4="relu",
        l6_act1="mish", l6_act2="relu6",
        l7_act1="identity", l7_act2=("biased_prelu", {"num_parameters": 1})),
}


def gen_projection_cases(mps):
    """Added after the first fixture set: own generator, so the files above keep their bytes."""
    g = torch.Generator().manual_seed(4321)
    for seed, (name, spec) in enumerate(PROJ_SPECS.items(), start=51):
        sd = O.make_pix_shuffle_state_dict(spec, seed)
        model = ref_pix_shuffle(mps, spec)
        model.load_state_dict(sd, strict=True)
        x = torch.rand((2, 3, 44, 60), generator=g) * 1.1
        with torch.no_grad():
            y = model(x)
        np.savez_compressed(os.path.join(GOLD, f"pix_shuffle_{name}.npz"), seed=seed, x=x.numpy(), y=y.numpy())
        os.chmod(os.path.join(GOLD, f"pix_shuffle_{name}.npz"), 0o644)
        yo = O.pix_shuffle_forward(sd, spec, x)
        print(f"pix_shuffle {name}: oracle-vs-reference max|d| = {(y - yo).abs().max().item():.3e}")


def gen_quantize_cases():
    """dataset_generator/quantize.py reduce_color_depth_and_dither(dithering_method='none') for every colour depth and
    util.py post_apply_resolution_style for every style, on one seeded image (own generator: earlier fixtures keep their bytes)."""
    from PIL import Image
    sys.path.insert(0, os.path.join(REF, "dataset_generator"))
    import quantize as ref_q   # noqa: E402  (numba / sklearn / PIL are present in this container)
    import util as ref_u       # noqa: E402
    rs = np.random.RandomState(77)
    img = rs.randint(0, 256, (22, 36, 3)).astype(np.uint8)
    img[0, :6] = [[0, 15, 16], [17, 31, 32], [239, 240, 255], [7, 8, 9], [3, 4, 5], [251, 252, 253]]   # grid boundaries
    out = {"img": img}
    for cs in ("RGB888", "RGB444", "RGB555", "RGB565", "RGB666"):
        q = ref_q.reduce_color_depth_and_dither(img, cs, dithering_method="none", verbose=0)
        out[f"q_{cs}"] = q
        print(f"quantize {cs}: oracle-vs-reference equal = {np.array_equal(q, O.quantize_grid(img, cs))}")
    q444 = out["q_RGB444"]
    for style in ref_u.SUPPORTED_RESOLUTION_STYLES:
        up = np.asarray(ref_u.post_apply_resolution_style(Image.fromarray(q444), style))
        out[f"post_{style}"] = up
        print(f"post style {style}: {up.shape}, oracle-vs-reference equal = {np.array_equal(up, O.post_resolution_style(q444, style))}")
    np.savez_compressed(os.path.join(GOLD, "quantize.npz"), **out)
    os.chmod(os.path.join(GOLD, "quantize.npz"), 0o644)


# kernel sizes other than 3 (model_pix_shuffle.py:21-64, 108-115) and the residual-UNet building block (residual_feature_block.py)
KSIZE_SPECS = {
    "ksize_a": O.pix_shuffle_preset("lightweight").with_kernels(5, 3, 5, 3, 7, 5, 3),
    "ksize_b": O.PixShuffleSpec((16, 16, 24, 24, 16, 16)).with_acts(l1_act1="mish", l2_act2=("biased_relu", {"num_parameters": 16}),
                                                                    l6_act2=("prelu", None), l7_act1="tanh",
                                                                    l7_act2="identity").with_kernels(7, 1, 3, 5, 1, 3, 5),
}
RESBLOCK_CASES = {   # name: (in, mid, out, kernel_size, acts)
    "resblock_a": (24, 16, 24, 3, {"act1": "identity", "act1_params": None, "act2": "relu", "act2_params": None,
                                   "act3": "identity", "act3_params": None, "act4": "relu", "act4_params": None}),
    "resblock_b": (20, 32, 40, 5, {"act1": "mish", "act1_params": None, "act2": "biased_prelu", "act2_params": {"num_parameters": 32},
                                   "act3": "sinlu", "act3_params": None, "act4": "prelu", "act4_params": {"num_parameters": 40}}),
    "resblock_c": (16, 16, 16, 7, {"act1": "telu", "act1_params": None, "act2": "relu6", "act2_params": None,
                                   "act3": "tanh", "act3_params": None, "act4": "leaky_relu", "act4_params": {"negative_slope": 0.1}}),
}


def gen_ksize_and_block_cases():
    mps, _, _, _ = import_reference()
    import residual_feature_block as ref_rfb   # noqa: E402  (imports only `activations`)
    g = torch.Generator().manual_seed(777)
    for seed, (name, spec) in enumerate(KSIZE_SPECS.items(), start=71):
        sd = O.make_pix_shuffle_state_dict(spec, seed)
        model = ref_pix_shuffle(mps, spec)
        model.load_state_dict(sd, strict=True)
        x = torch.rand((2, 3, 44, 60), generator=g) * 1.1
        with torch.no_grad():
            y = model(x)
        np.savez_compressed(os.path.join(GOLD, f"pix_shuffle_{name}.npz"), seed=seed, x=x.numpy(), y=y.numpy())
        print(f"pix_shuffle {name}: oracle-vs-reference max|d| = {(y - O.pix_shuffle_forward(sd, spec, x)).abs().max().item():.3e}")
    for seed, (name, (ci, cm, co, ks, acts)) in enumerate(RESBLOCK_CASES.items(), start=81):
        sd = O.make_residual_block_state_dict(ci, cm, co, ks, acts, seed)
        blk = ref_rfb.ResidualFeatureBlock(ci, cm, co, ks, acts=acts).eval()
        blk.load_state_dict(sd, strict=True)
        x = torch.randn((2, ci, 30, 44), generator=g) * 0.7
        with torch.no_grad():
            y = blk(x)
        np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), seed=seed, x=x.numpy(), y=y.numpy())
        print(f"{name}: oracle-vs-reference max|d| = {(y - O.residual_block_forward(sd, acts, x)).abs().max().item():.3e}")
    for f in os.listdir(GOLD):
        if f.endswith(".npz"):
            os.chmod(os.path.join(GOLD, f), 0o644)


def gen_dither_cases():
    """dataset_generator/quantize.py: the numba checkerboard / ordered dither kernels and the palette mapping, called
    directly on seeded images with fixed palettes (palette generation -- k-means etc. -- is not part of the kernels)."""
    sys.path.insert(0, os.path.join(REF, "dataset_generator"))
    import quantize as ref_q   # noqa: E402
    rs = np.random.RandomState(99)
    h, w = 37, 53
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 255 // (w - 1)), (yy * 255 // (h - 1)), ((xx + yy) * 255 // (h + w - 2))], axis=2).astype(np.uint8)
    img[::3, ::2] = rs.randint(0, 256, img[::3, ::2].shape)                      # gradient + noise
    out = {"img": img}
    pals = {"p2": np.array([[0, 0, 0], [255, 255, 255]], np.uint8),
            "p16": (rs.randint(0, 16, (16, 3)) * 17).astype(np.uint8),
            "p64dup": np.repeat((rs.randint(0, 16, (32, 3)) * 16).astype(np.uint8), 2, axis=0),   # duplicate colours: tie-breaks
            "p1": np.array([[10, 200, 30]], np.uint8)}
    img[0, :16] = pals["p16"]                                                  # exact palette hits stay undithered
    for pname, pal in pals.items():
        out[pname] = pal
        palf = pal.astype(np.float64)
        imf = img.astype(np.float64)
        res = np.zeros_like(img)
        ref_q._apply_checkerboard_dithering_numba_optimized(imf, palf, pal, res)
        out[f"{pname}_checkerboard"] = res.copy()
        for name, mat in (("bayer2x2", ref_q.BAYER_MATRIX_2X2), ("bayer4x4", ref_q.BAYER_MATRIX_4X4), ("bayer8x8", ref_q.BAYER_MATRIX_8X8)):
            res = np.zeros_like(img)
            ref_q._apply_ordered_dithering_numba_optimized(imf, palf, pal, res, mat.astype(np.float64) / (mat.shape[0] ** 2))
            out[f"{pname}_{name}"] = res.copy()
        dist = np.sum((imf.reshape(-1, 3)[:, np.newaxis, :] - palf) ** 2, axis=2)      # quantize.py:533-536
        out[f"{pname}_none"] = pal[np.argmin(dist, axis=1)].reshape(img.shape)
        for m in ("none", "checkerboard", "bayer2x2", "bayer4x4", "bayer8x8"):
            print(f"dither {pname} {m}: oracle-vs-reference equal = {np.array_equal(out[f'{pname}_{m}'], O.dither_palette(img, pal, m))}")
    np.savez_compressed(os.path.join(GOLD, "dither.npz"), **out)
    os.chmod(os.path.join(GOLD, "dither.npz"), 0o644)


def gen_round2_fixtures():
    """Added in round 2 (own option: the earlier fixtures keep their bytes): the other five screenshots + shipped
    predictions of pix_shuffle, conv3_heavy's trained weights with two of its shipped predictions, and the names of the
    `perceptual_criterion.*` state_dict entries a genuine reference checkpoint carries (train.py:236/246 saves
    model.state_dict(); every reference Model owns a PerceptualLoss with a torchvision VGG16, loss_vgg.py:60)."""
    for i in range(8):
        for src, dst in ((f"model/samples/sample{i}.png", "samples"),
                         (f"model/model_pix_shuffle/predicted/sample{i}.png", "predicted_pix_shuffle")):
            os.makedirs(os.path.join(GOLD, dst), exist_ok=True)
            if not os.path.exists(os.path.join(GOLD, dst, f"sample{i}.png")):
                shutil.copy(os.path.join(REF, src), os.path.join(GOLD, dst))
    os.makedirs(os.path.join(GOLD, "predicted_conv3_heavy"), exist_ok=True)
    for i in (5, 6):
        shutil.copy(os.path.join(REF, f"model/model_conv3_heavy/predicted/sample{i}.png"), os.path.join(GOLD, "predicted_conv3_heavy"))
        shutil.copy(os.path.join(REF, f"model/model_conv3/predicted/sample{i}.png"), os.path.join(GOLD, "predicted_conv3"))
    sd3 = onnx_weights.conv3_state_dict_from_onnx(os.path.join(REF, "model/model_conv3_heavy/conv3_heavy.onnx"))
    np.savez_compressed(os.path.join(GOLD, "conv3_heavy_trained_fp16.npz"),
                        **{k: (v.numpy().astype(np.float16) if v.dtype.is_floating_point else v.numpy()) for k, v in sd3.items()})
    # key names of the loss module inside a reference checkpoint: PerceptualLoss.vgg = torchvision vgg16 (loss_vgg.py:60);
    # torchvision is present here, the ImageNet weights are not needed for the names
    import torchvision
    vgg = torchvision.models.vgg16(weights=None)
    keys = ["perceptual_criterion.vgg." + k for k in vgg.state_dict()]
    shapes = [list(v.shape) for v in vgg.state_dict().values()]
    import json
    with open(os.path.join(GOLD, "reference_checkpoint_loss_keys.json"), "w") as f:
        json.dump({"keys": keys, "shapes": shapes}, f)
    for root, _, files in os.walk(GOLD):
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    print(f"round-2 fixtures written: 8 screenshots, conv3_heavy weights, {len(keys)} loss-module keys")


def main():
    if "--only-round2" in sys.argv:
        gen_round2_fixtures()
        return
    if "--only-dither" in sys.argv:
        gen_dither_cases()
        return
    if "--only-ksize" in sys.argv:
        gen_ksize_and_block_cases()
        return
    if "--only-quantize" in sys.argv:
        gen_quantize_cases()
        return
    mps, mc3, mc5, gamma = import_reference()
    if "--only-projections" in sys.argv:
        gen_projection_cases(mps)
        return
    if os.path.isdir(GOLD):
        shutil.rmtree(GOLD)
    os.makedirs(GOLD)
    torch.set_num_threads(8)
    g = torch.Generator().manual_seed(1234)

    # ---- pix_shuffle, float path (model_pix_shuffle.py:227-298) ----
    cases = {"lightweight": O.pix_shuffle_preset("lightweight"),
             "heavyweight": O.pix_shuffle_preset("heavyweight"), **VOCAB_SPECS}
    for seed, (name, spec) in enumerate(cases.items(), start=11):
        sd = O.make_pix_shuffle_state_dict(spec, seed)
        model = ref_pix_shuffle(mps, spec)
        missing = model.load_state_dict(sd, strict=True)
        # inputs: uniform rand (the reference's own benchmark input, :348) incl. a frame that
        # exceeds [0,1] a little, odd tile-unfriendly sizes, batch 2
        x = torch.rand((2, 3, 44, 60), generator=g) * 1.1
        with torch.no_grad():
            y = model(x)
        np.savez_compressed(os.path.join(GOLD, f"pix_shuffle_{name}.npz"), seed=seed,
                            x=x.numpy(), y=y.numpy())
        yo = O.pix_shuffle_forward(sd, spec, x)
        print(f"pix_shuffle {name}: oracle-vs-reference max|d| = {(y - yo).abs().max().item():.3e}")

    # ---- conv3 / conv5 (model_conv3.py:102-155, model_conv5.py:114-151) ----
    for preset in ("lightweight", "heavyweight"):
        seed = 31 if preset == "lightweight" else 32
        sd = O.make_bn_state_dict(O.conv3_channels(preset), seed)
        m = mc3.get_model(preset).eval()
        m.load_state_dict(sd, strict=True)
        xu = torch.randint(0, 256, (2, 4, 36, 52), generator=g, dtype=torch.uint8)
        with torch.no_grad():
            y = m(xu)
        np.savez_compressed(os.path.join(GOLD, f"conv3_{preset}.npz"), seed=seed, x=xu.numpy(), y=y.numpy())
        print(f"conv3 {preset}: max|d| = {(y - O.conv3_forward(sd, xu)).abs().max().item():.3e}")

        seed += 10
        sd = O.make_bn_state_dict(O.conv5_channels(preset), seed)
        m = mc5.get_model(preset).eval()
        m.load_state_dict(sd, strict=True)
        x = torch.rand((2, 3, 36, 52), generator=g)
        with torch.no_grad():
            y = m(x)
        np.savez_compressed(os.path.join(GOLD, f"conv5_{preset}.npz"), seed=seed, x=x.numpy(), y=y.numpy())
        print(f"conv5 {preset}: max|d| = {(y - O.conv5_forward(sd, x)).abs().max().item():.3e}")

    # ---- gamma (gamma.py:13-15, 31-33) ----
    t = torch.linspace(0, 1.3, 257)
    np.savez_compressed(os.path.join(GOLD, "gamma.npz"), t=t.numpy(),
                        to_linear=gamma.srgb_to_linear_approx(t).numpy(),
                        to_srgb=gamma.linear_to_srgb_approx(t).numpy())

    # ---- trained weights + the reference's own regression artefacts ----
    sd = onnx_weights.pix_shuffle_state_dict_from_onnx(os.path.join(REF, "model/model_pix_shuffle/pix_shuffle.onnx"))
    np.savez_compressed(os.path.join(GOLD, "pix_shuffle_trained_fp16.npz"),
                        **{k: v.numpy().astype(np.float16) for k, v in sd.items()})
    for preset, d in (("lightweight", "model_conv3"),):
        sd3 = onnx_weights.conv3_state_dict_from_onnx(os.path.join(REF, f"model/{d}/conv3.onnx"))
        np.savez_compressed(os.path.join(GOLD, "conv3_trained_fp16.npz"),
                            **{k: (v.numpy().astype(np.float16) if v.dtype.is_floating_point else v.numpy())
                               for k, v in sd3.items()})
    os.makedirs(os.path.join(GOLD, "samples"))
    os.makedirs(os.path.join(GOLD, "predicted_pix_shuffle"))
    os.makedirs(os.path.join(GOLD, "predicted_conv3"))
    for i in (1, 5, 6):   # three of the eight screenshots keep the fixture set small
        shutil.copy(os.path.join(REF, f"model/samples/sample{i}.png"), os.path.join(GOLD, "samples"))
        shutil.copy(os.path.join(REF, f"model/model_pix_shuffle/predicted/sample{i}.png"),
                    os.path.join(GOLD, "predicted_pix_shuffle"))
        shutil.copy(os.path.join(REF, f"model/model_conv3/predicted/sample{i}.png"),
                    os.path.join(GOLD, "predicted_conv3"))
    gen_projection_cases(mps)
    gen_quantize_cases()
    gen_round2_fixtures()
    gen_dither_cases()
    gen_ksize_and_block_cases()
    for root, _, files in os.walk(GOLD):
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
