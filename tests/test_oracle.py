"""CPU: pin the oracle (oracle/enhancer_oracle.py) against vectors produced by the REAL reference
(oracle/gen_golden.py) and against the reference's own shipped regression artefacts."""
import os

import numpy as np
import pytest
import torch

from oracle import enhancer_oracle as O
from tests.util import GOLD, gold_spec, load_gold, load_png_rgb, load_png_rgba, trained_conv3_sd, trained_pix_shuffle_sd


@pytest.mark.parametrize("name", ["lightweight", "heavyweight", "vocab_a", "vocab_b", "proj_a"])
def test_pix_shuffle_matches_reference_vectors(name):
    g = load_gold(f"pix_shuffle_{name}")
    spec = gold_spec(name)
    sd = O.make_pix_shuffle_state_dict(spec, int(g["seed"]))
    y = O.pix_shuffle_forward(sd, spec, torch.from_numpy(g["x"]))
    assert y.shape == g["y"].shape
    assert np.abs(y.numpy() - g["y"]).max() <= 2e-6     # same torch ops; tiny slack for thread count


@pytest.mark.parametrize("preset", ["lightweight", "heavyweight"])
def test_conv3_conv5_match_reference_vectors(preset):
    g = load_gold(f"conv3_{preset}")
    sd = O.make_bn_state_dict(O.conv3_channels(preset), int(g["seed"]))
    y = O.conv3_forward(sd, torch.from_numpy(g["x"]))
    assert np.abs(y.numpy() - g["y"]).max() <= 2e-3      # 0..255 scale
    assert (y[:, 3] == 255.0).all()
    g = load_gold(f"conv5_{preset}")
    sd = O.make_bn_state_dict(O.conv5_channels(preset), int(g["seed"]))
    y = O.conv5_forward(sd, torch.from_numpy(g["x"]))
    assert np.abs(y.numpy() - g["y"]).max() <= 2e-6


def test_conv3_rejects_non_uint8_rgba():
    sd = O.make_bn_state_dict(O.conv3_channels("lightweight"), 1)
    with pytest.raises(ValueError):
        O.conv3_forward(sd, torch.zeros(1, 3, 8, 8, dtype=torch.uint8))
    with pytest.raises(ValueError):
        O.conv3_forward(sd, torch.zeros(1, 4, 8, 8))


def test_gamma_matches_reference():
    g = load_gold("gamma")
    t = torch.from_numpy(g["t"])
    assert np.abs(O.gamma_in(t).numpy() - g["to_linear"]).max() <= 1e-7
    assert np.abs(O.gamma_out(t).numpy() - g["to_srgb"]).max() <= 1e-7


@pytest.mark.parametrize("i", [1, 5, 6])
def test_trained_weights_reproduce_shipped_screenshots(i):
    """ONNX-embedded trained weights + model/samples/sample{i}.png -> the shipped
    model/model_pix_shuffle/predicted/sample{i}.png through the float glue of train.py:57-73."""
    sd = trained_pix_shuffle_sd()
    spec = O.pix_shuffle_preset("lightweight")
    x = load_png_rgb(os.path.join(GOLD, "samples", f"sample{i}.png"))
    want = load_png_rgb(os.path.join(GOLD, "predicted_pix_shuffle", f"sample{i}.png"))
    got = O.float_pipeline(sd, spec, x)
    assert got.shape == want.shape == (1, 3, 576, 752)
    d = (got.int() - want.int()).abs()
    assert O.psnr(got, want, 255.0) >= 60.0
    assert d.max().item() <= 4
    assert (d == 0).float().mean().item() >= 0.98


@pytest.mark.parametrize("i", [1, 5, 6])
def test_trained_conv3_reproduces_shipped_screenshots(i):
    sd = trained_conv3_sd()
    x = load_png_rgba(os.path.join(GOLD, "samples", f"sample{i}.png")).permute(0, 3, 1, 2).contiguous()
    want = load_png_rgba(os.path.join(GOLD, "predicted_conv3", f"sample{i}.png")).permute(0, 3, 1, 2)
    got = O.conv3_forward(sd, x).clamp(0, 255).to(torch.uint8)
    d = (got.int() - want.int()).abs()
    assert d.max().item() <= 1
    assert O.psnr(got, want, 255.0) >= 58.0


def test_framebuffer_contract_shape_alpha_and_crop():
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 3)
    fb = O.synth_framebuffers(2, seed=5, h=32, w=48)
    out = O.framebuffer_forward(sd, spec, fb)
    assert out.shape == fb.shape and out.dtype == torch.uint8 and (out[..., 3] == 255).all()
    outc = O.framebuffer_forward(sd, spec, fb, crop16=True)
    assert (outc[:, :, :16, :3] == 0).all() and (outc[..., 3] == 255).all()


def test_synth_framebuffers_are_rgb444_with_pixel_modes():
    fb = O.synth_framebuffers(4, seed=0, h=16, w=24)
    assert (fb[..., :3] % 17 == 0).all() and (fb[..., 3] == 255).all()
    lores, lores_laced, hires, hires_laced = fb[0], fb[1], fb[2], fb[3]
    assert (lores[0::2] == lores[1::2]).all() and (lores[:, 0::2] == lores[:, 1::2]).all()
    assert (lores_laced[:, 0::2] == lores_laced[:, 1::2]).all() and not (lores_laced[0::2] == lores_laced[1::2]).all()
    assert (hires[0::2] == hires[1::2]).all() and not (hires[:, 0::2] == hires[:, 1::2]).all()
    assert not (hires_laced[0::2] == hires_laced[1::2]).all()


def test_quantize_and_resolution_style_match_reference_vectors():
    """oracle quantize_grid / post_resolution_style == dataset_generator/quantize.py:512-521 (dithering 'none') and
    util.py:318-350, on the vectors gen_golden.py produced by calling the reference's own functions."""
    g = load_gold("quantize")
    for cs in ("RGB888", "RGB444", "RGB555", "RGB565", "RGB666"):
        assert np.array_equal(O.quantize_grid(g["img"], cs), g[f"q_{cs}"])
    for style in ("lores", "lores_laced", "hires", "hires_laced"):
        assert np.array_equal(O.post_resolution_style(g["q_RGB444"], style), g[f"post_{style}"])
    with pytest.raises(ValueError):
        O.quantize_grid(g["img"], "RGB333")


def test_synthetic_stream_is_rgb444_in_the_four_pixel_modes():
    fr = O.synth_rgb444_frames(8, 16, 24, seed=5)
    assert fr.shape == (8, 16, 24, 4) and (fr[..., 3] == 255).all() and (fr[..., :3] % 17 == 0).all()
    for i, style in enumerate(O.SYNTH_STYLE_OF_FRAME * 2):
        sy, sx = O.PIXEL_MODES[style]
        cells = fr[i, ::sy, ::sx]
        assert np.array_equal(np.repeat(np.repeat(cells, sy, 0), sx, 1), fr[i])        # sy x sx blocks
    assert np.array_equal(O.synth_rgb444_frames(3, 16, 24, seed=5, first_frame=4), fr[4:7])   # a stream can be cut anywhere
    assert not np.array_equal(fr[0], fr[4]) and len(np.unique(fr[3, :, :, 0])) == 16
    assert (O.synth_rgb444_frames(1, 8, 8, seed=5, expand17=False)[..., :3] % 16 == 0).all()


def test_dither_oracle_matches_the_reference_numba_kernels():
    """tests/golden/dither.npz holds outputs of the reference's own _apply_checkerboard_dithering_numba_optimized /
    _apply_ordered_dithering_numba_optimized (quantize.py:137-331) and its palette mapping (:529-537)."""
    g = load_gold("dither")
    for pname in ("p2", "p16", "p64dup", "p1"):
        for method in ("none", "checkerboard", "bayer2x2", "bayer4x4", "bayer8x8"):
            assert np.array_equal(O.dither_palette(g["img"], g[pname], method), g[f"{pname}_{method}"]), (pname, method)
    assert (O.dither_palette(g["img"], np.zeros((0, 3), np.uint8), "checkerboard") == 0).all()


@pytest.mark.parametrize("name", ["ksize_a", "ksize_b"])
def test_oracle_kernel_sizes_match_reference_vectors(name):
    """layer{i}_kernel_size in {1, 3, 5, 7} (model_pix_shuffle.py:21-64, 108-115): vectors from the real reference Model."""
    g = load_gold(f"pix_shuffle_{name}")
    spec = gold_spec(name)
    sd = O.make_pix_shuffle_state_dict(spec, int(g["seed"]))
    y = O.pix_shuffle_forward(sd, spec, torch.from_numpy(g["x"])).numpy()
    assert np.abs(y - g["y"]).max() <= 1e-6


@pytest.mark.parametrize("name", ["resblock_a", "resblock_b", "resblock_c"])
def test_oracle_residual_feature_block_matches_reference_vectors(name):
    """residual_feature_block.py:5-55 (1x1 -> k x k -> 1x1 bottleneck, optional projection): vectors from the real reference block."""
    from oracle.gen_golden import RESBLOCK_CASES
    ci, cm, co, ks, acts = RESBLOCK_CASES[name]
    g = load_gold(name)
    sd = O.make_residual_block_state_dict(ci, cm, co, ks, acts, int(g["seed"]))
    y = O.residual_block_forward(sd, acts, torch.from_numpy(g["x"])).numpy()
    assert np.abs(y - g["y"]).max() <= 1e-6
