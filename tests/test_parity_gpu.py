"""GPU: the engine (through the drop-in modules -> ctypes -> C ABI -> CUDA) against the oracle on the
same seeded inputs, against the committed reference vectors, and -- at full 752x576 size -- through
size-independent properties.  Tolerances are SURVEY.md section 8d / BASELINE.md section 4."""
import os

import numpy as np
import pytest
import torch

from oracle import enhancer_oracle as O
from tests.util import (GOLD, build_pkg_pix_shuffle, gold_spec, load_gold, load_png_rgb, load_png_rgba,
                        trained_conv3_sd, trained_pix_shuffle_sd)

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5          # fp32 build vs fp32 oracle, float output in [0,~1.2]
FP32_TOL_255 = 2e-3      # conv3 output scale 0..255
BF16_TOL = 1e-2          # bf16 build max-abs
BF16_PSNR = 55.0         # bf16 build PSNR (peak 1.0)


def dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return torch.device("cuda", 0)


def test_native_library_is_what_runs():
    from fs_uae_image_enhancer_project_b200 import _lib
    spec = O.pix_shuffle_preset("lightweight")
    m = build_pkg_pix_shuffle(spec, O.make_pix_shuffle_state_dict(spec, 1)).to(dev())
    y = m(torch.rand(1, 3, 16, 16, device=dev()))
    assert y.shape == (1, 3, 16, 16)
    eng = m.engine_for(dev(), 16, 16)
    assert eng.variant == "fp32_fma" and eng.last_launch_count >= 9
    assert any("libfsuae_enhancer.so" in l for l in open("/proc/self/maps"))


@pytest.mark.parametrize("name", ["lightweight", "heavyweight", "vocab_a", "vocab_b", "proj_a"])
def test_fp32_pix_shuffle_matches_reference_vectors(name):
    g = load_gold(f"pix_shuffle_{name}")
    spec = gold_spec(name)
    sd = O.make_pix_shuffle_state_dict(spec, int(g["seed"]))
    m = build_pkg_pix_shuffle(spec, sd).to(dev())
    y = m(torch.from_numpy(g["x"]).to(dev())).cpu().numpy()
    assert np.abs(y - g["y"]).max() <= FP32_TOL


@pytest.mark.parametrize("preset", ["lightweight", "heavyweight"])
def test_fp32_conv3_conv5_match_reference_vectors(preset):
    from fs_uae_image_enhancer_project_b200 import model_conv3, model_conv5
    g = load_gold(f"conv3_{preset}")
    m = model_conv3.get_model(preset)
    m.load_state_dict(O.make_bn_state_dict(O.conv3_channels(preset), int(g["seed"])))
    y = m.to(dev())(torch.from_numpy(g["x"]).to(dev())).cpu().numpy()
    assert y.shape == g["y"].shape and np.abs(y - g["y"]).max() <= FP32_TOL_255
    assert (y[:, 3] == 255.0).all()
    g = load_gold(f"conv5_{preset}")
    m = model_conv5.get_model(preset)
    m.load_state_dict(O.make_bn_state_dict(O.conv5_channels(preset), int(g["seed"])))
    y = m.to(dev())(torch.from_numpy(g["x"]).to(dev())).cpu().numpy()
    assert np.abs(y - g["y"]).max() <= FP32_TOL


@pytest.mark.parametrize("shape", [(1, 2, 2), (1, 2, 66), (3, 34, 2), (2, 18, 70), (1, 64, 96), (5, 6, 130)])
def test_fp32_edge_shapes_and_borders(shape):
    """Tiny / ragged frames: every pixel is a border pixel somewhere; zero padding is per layer."""
    B, H, W = shape
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 21)
    x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(H * W))
    want = O.pix_shuffle_forward(sd, spec, x)
    got = build_pkg_pix_shuffle(spec, sd).to(dev())(x.to(dev())).cpu()
    assert (got - want).abs().max().item() <= FP32_TOL


def test_fp32_batch_larger_than_chunk_and_empty_batch():
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 22)
    m = build_pkg_pix_shuffle(spec, sd).to(dev())
    m.chunk_frames = 3
    x = torch.rand(8, 3, 20, 28, generator=torch.Generator().manual_seed(3))
    got = m(x.to(dev())).cpu()
    assert (got - O.pix_shuffle_forward(sd, spec, x)).abs().max().item() <= FP32_TOL
    assert m(torch.zeros(0, 3, 20, 28, device=dev())).shape == (0, 3, 20, 28)


def test_forward_is_stateless_and_rebuilds_after_load_state_dict():
    spec = O.pix_shuffle_preset("lightweight")
    sd1, sd2 = O.make_pix_shuffle_state_dict(spec, 1), O.make_pix_shuffle_state_dict(spec, 2)
    m = build_pkg_pix_shuffle(spec, sd1).to(dev())
    x = torch.rand(1, 3, 24, 24, generator=torch.Generator().manual_seed(9))
    a = m(x.to(dev())).cpu()
    b = m(x.to(dev())).cpu()
    assert torch.equal(a, b)
    m.load_state_dict(sd2)
    c = m(x.to(dev())).cpu()
    assert (c - O.pix_shuffle_forward(sd2, spec, x)).abs().max().item() <= FP32_TOL
    assert (c - a).abs().max().item() > 1e-3


def test_error_behaviour_on_gpu():
    from fs_uae_image_enhancer_project_b200 import model_conv3
    spec = O.pix_shuffle_preset("lightweight")
    m = build_pkg_pix_shuffle(spec, O.make_pix_shuffle_state_dict(spec, 1)).to(dev())
    with pytest.raises(ValueError, match="even"):
        m(torch.zeros(1, 3, 15, 16, device=dev()))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 4, 16, 16, device=dev()))
    m3 = model_conv3.get_model("lightweight").to(dev())
    with pytest.raises(ValueError, match="uint8"):
        m3(torch.zeros(1, 3, 16, 16, device=dev()))


@pytest.mark.parametrize("crop16", [False, True])
def test_fp32_framebuffer_contract_full_size(crop16):
    """Deployed uint8 RGBA contract at 752x576 on mixed pixel-mode RGB444 frames: <= 1 LSB."""
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 31)
    fb = O.synth_framebuffers(4, seed=8)
    want = O.framebuffer_forward(sd, spec, fb, crop16=crop16)
    m = build_pkg_pix_shuffle(spec, sd).to(dev())
    got = m.forward_framebuffer(fb.to(dev()), crop16=crop16).cpu()
    d = (got.int() - want.int()).abs()
    assert d.max().item() <= 1 and (d == 0).float().mean().item() >= 0.999
    assert (got[..., 3] == 255).all()
    if crop16:
        assert (got[:, :, :16, :3] == 0).all()
    # host-buffer entry point gives the identical bytes
    host = m.run_host(fb.pin_memory(), crop16=crop16)
    assert torch.equal(host, got)


@pytest.mark.parametrize("i", [1, 5, 6])
def test_fp32_trained_weights_reproduce_shipped_screenshots(i):
    sd = trained_pix_shuffle_sd()
    spec = O.pix_shuffle_preset("lightweight")
    m = build_pkg_pix_shuffle(spec, sd).to(dev())
    rgba = load_png_rgba(os.path.join(GOLD, "samples", f"sample{i}.png"))
    want = load_png_rgb(os.path.join(GOLD, "predicted_pix_shuffle", f"sample{i}.png"))
    got = m.forward_framebuffer(rgba.to(dev())).cpu()[..., :3].permute(0, 3, 1, 2)
    assert O.psnr(got, want, 255.0) >= 60.0
    assert (got.int() - want.int()).abs().max().item() <= 4


def test_fp32_trained_conv3_reproduces_shipped_screenshots():
    from fs_uae_image_enhancer_project_b200 import model_conv3
    m = model_conv3.get_model("lightweight")
    m.load_state_dict(trained_conv3_sd())
    m = m.to(dev())
    x = load_png_rgba(os.path.join(GOLD, "samples", "sample5.png")).permute(0, 3, 1, 2).contiguous()
    want = load_png_rgba(os.path.join(GOLD, "predicted_conv3", "sample5.png")).permute(0, 3, 1, 2)
    got = m(x.to(dev())).cpu().clamp(0, 255).to(torch.uint8)
    assert (got.int() - want.int()).abs().max().item() <= 1


def test_fp32_full_size_properties():
    """752x576: (a) frames are independent -- a batch equals its frames run one by one, in any
    order; (b) translation equivariance away from the border: shifting the input by (2,2) full-res
    pixels shifts the interior of the output; (c) output >= 0 (final ReLU)."""
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 41)
    m = build_pkg_pix_shuffle(spec, sd).to(dev())
    x = torch.rand(3, 3, 576, 752, generator=torch.Generator().manual_seed(1)).to(dev())
    y = m(x)
    assert (y >= 0).all()
    y_rev = m(x.flip(0).contiguous()).flip(0)
    assert torch.equal(y, y_rev)
    y1 = m(x[1:2].contiguous())
    assert torch.equal(y[1:2], y1)
    xs = torch.roll(x[:1], shifts=(2, 2), dims=(2, 3)).contiguous()
    ys = m(xs)
    a = ys[:, :, 32:-32, 32:-32]
    b = torch.roll(y[:1], shifts=(2, 2), dims=(2, 3))[:, :, 32:-32, 32:-32]
    assert (a - b).abs().max().item() <= 1e-6
    # spot-check against the oracle on one full-size frame
    want = O.pix_shuffle_forward(sd, spec, x[:1].cpu())
    assert (y[:1].cpu() - want).abs().max().item() <= FP32_TOL


# ----------------------------------------------------------------------------------------------
# bf16 tensor-core build (tcgen05): tolerance max-abs <= 1e-2, PSNR >= 55 dB vs the fp32 oracle
# ----------------------------------------------------------------------------------------------

def _bf16_model(spec, sd):
    return build_pkg_pix_shuffle(spec, sd).to(dev()).set_precision("bf16")


@pytest.mark.parametrize("shape", [(1, 16, 16), (2, 64, 96), (1, 40, 300), (3, 34, 254), (1, 6, 508), (2, 2, 2)])
def test_bf16_pix_shuffle_small_frames(shape):
    """Single strip, exactly-two-strips (W/2 = 127), three strips, 1-pixel-high maps."""
    B, H, W = shape
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 51)
    x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(H + W))
    want = O.pix_shuffle_forward(sd, spec, x)
    m = _bf16_model(spec, sd)
    got = m(x.to(dev())).cpu()
    assert m.engine_for(dev(), H, W).variant == "bf16_tcgen05"
    assert (got - want).abs().max().item() <= BF16_TOL
    assert O.psnr(got, want, 1.0) >= BF16_PSNR


def test_bf16_matches_reference_vectors_and_trained_weights():
    g = load_gold("pix_shuffle_lightweight")
    spec = gold_spec("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, int(g["seed"]))
    got = _bf16_model(spec, sd)(torch.from_numpy(g["x"]).to(dev())).cpu().numpy()
    assert np.abs(got - g["y"]).max() <= BF16_TOL
    # trained weights on a real Amiga screenshot vs the shipped prediction (u8, <= 6 LSB, PSNR >= 50 dB)
    m = _bf16_model(spec, trained_pix_shuffle_sd())
    rgba = load_png_rgba(os.path.join(GOLD, "samples", "sample5.png"))
    want = load_png_rgb(os.path.join(GOLD, "predicted_pix_shuffle", "sample5.png"))
    out = m.forward_framebuffer(rgba.to(dev())).cpu()[..., :3].permute(0, 3, 1, 2)
    d = (out.int() - want.int()).abs()
    print(f"bf16 trained sample5: max {d.max().item()} LSB, psnr {O.psnr(out, want, 255.0):.1f} dB, "
          f"<=1 LSB {(d <= 1).float().mean().item():.4f}")
    # bf16 rounding of near-black linear values is amplified by the 1/2.2 gamma (slope ~13 at L=0.002)
    assert d.max().item() <= 16
    assert O.psnr(out, want, 255.0) >= 48.0
    assert (d <= 1).float().mean().item() >= 0.97


@pytest.mark.parametrize("crop16", [False, True])
def test_bf16_framebuffer_contract_full_size(crop16):
    """uint8 RGBA end to end at 752x576 on mixed pixel-mode frames: <= 6 LSB, >= 85 % exact."""
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 31)
    fb = O.synth_framebuffers(4, seed=8)
    want = O.framebuffer_forward(sd, spec, fb, crop16=crop16)
    m = _bf16_model(spec, sd)
    got = m.forward_framebuffer(fb.to(dev()), crop16=crop16).cpu()
    d = (got.int() - want.int()).abs()
    assert d.max().item() <= 6 and (d == 0).float().mean().item() >= 0.85
    assert (got[..., 3] == 255).all()
    if crop16:
        assert (got[:, :, :16, :3] == 0).all()
    assert torch.equal(m.run_host(fb.pin_memory(), crop16=crop16), got)


def test_bf16_full_size_float_and_batch_properties():
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 41)
    m = _bf16_model(spec, sd)
    m.chunk_frames = 4
    x = torch.rand(6, 3, 576, 752, generator=torch.Generator().manual_seed(2)).to(dev())
    y = m(x)
    assert (y >= 0).all()
    assert torch.equal(y[4:5], m(x[4:5].contiguous()))          # frames are independent, chunking is invisible
    assert torch.equal(y.flip(0), m(x.flip(0).contiguous()))
    want = O.pix_shuffle_forward(sd, spec, x[:2].cpu())
    got = y[:2].cpu()
    assert (got - want).abs().max().item() <= BF16_TOL
    assert O.psnr(got, want, 1.0) >= BF16_PSNR


def test_bf16_unsupported_network_fails_loudly():
    """A channel softmax in the last layer (it would have to run inside the PixelShuffle tail) is not implemented on
    the bf16 build: explicit error, no fallback."""
    from fs_uae_image_enhancer_project_b200 import _lib
    spec = O.pix_shuffle_preset("lightweight").with_acts(l7_act1="softmax")
    mb = _bf16_model(spec, O.make_pix_shuffle_state_dict(spec, 1))
    with pytest.raises(_lib.EngineError) as ei:
        mb(torch.rand(1, 3, 16, 16, device=dev()))
    assert ei.value.code == _lib.ERR_UNSUPPORTED


def test_bf16_channel_softmax_slots():
    """vocab_b: softmax behind the residual add of layer 2 (wide kernel, residual from global memory) and log_softmax as
    the first slot of layer 3 (resident-weight kernel): three passes over the accumulator row in the run-time epilogue."""
    g = load_gold("pix_shuffle_vocab_b")
    spec = gold_spec("vocab_b")
    sd = O.make_pix_shuffle_state_dict(spec, int(g["seed"]))
    m = _bf16_model(spec, sd)
    got = m(torch.from_numpy(g["x"]).to(dev())).cpu()
    want = torch.from_numpy(g["y"])
    assert (got - want).abs().max().item() <= 2 * BF16_TOL and O.psnr(got, want, 1.0) >= 50.0   # log_softmax amplifies bf16 rounding
    x = torch.rand(2, 3, 40, 300, generator=torch.Generator().manual_seed(3))
    got = m(x.to(dev())).cpu()
    want = O.pix_shuffle_forward(sd, spec, x)
    assert (got - want).abs().max().item() <= 2 * BF16_TOL and O.psnr(got, want, 1.0) >= 50.0


def test_bf16_conv3_heavyweight_wide_kernel():
    """192 -> 256 -> 3 channels: neither the weights (884 KB) nor full-depth input rows fit in shared memory; the
    K-streamed tile kernel (CTA pairs, 8-row tiles, 2 output-channel groups) runs conv2 and conv3."""
    from fs_uae_image_enhancer_project_b200 import model_conv3
    g = load_gold("conv3_heavyweight")
    sd = O.make_bn_state_dict(O.conv3_channels("heavyweight"), int(g["seed"]))
    m = model_conv3.get_model("heavyweight")
    m.load_state_dict(sd)
    m = m.to(dev()).set_precision("bf16")
    y = m(torch.from_numpy(g["x"]).to(dev())).float().cpu()
    want = torch.from_numpy(g["y"])
    assert y.shape == want.shape and (y[:, 3] == 255.0).all()
    assert (y - want).abs().max().item() <= 255 * BF16_TOL and O.psnr(y, want, 255.0) >= 50.0
    # ragged geometry: 3 frames, 21 rows (tiles hang over the bottom edge), 300 columns (3 strips, last one partial)
    x = torch.randint(0, 256, (3, 4, 21, 300), dtype=torch.uint8, generator=torch.Generator().manual_seed(9))
    want = O.conv3_forward(sd, x)
    got = m(x.to(dev())).float().cpu()
    assert (got - want).abs().max().item() <= 255 * BF16_TOL and O.psnr(got, want, 255.0) >= 50.0
    assert torch.equal(got[1:2], m(x[1:2].contiguous().to(dev())).float().cpu())     # tiling is invisible


def test_bf16_pix_shuffle_heavyweight_matches_reference_vectors():
    """108-channel conv4/conv5 stream K through the wide tile kernel (packed weights would not fit otherwise)."""
    g = load_gold("pix_shuffle_heavyweight")
    spec = gold_spec("heavyweight")
    sd = O.make_pix_shuffle_state_dict(spec, int(g["seed"]))
    got = _bf16_model(spec, sd)(torch.from_numpy(g["x"]).to(dev())).cpu()
    want = torch.from_numpy(g["y"])
    assert (got - want).abs().max().item() <= BF16_TOL and O.psnr(got, want, 1.0) >= BF16_PSNR
    x = torch.rand(1, 3, 64, 600, generator=torch.Generator().manual_seed(4))      # three strips
    got = _bf16_model(spec, sd)(x.to(dev())).cpu()
    want = O.pix_shuffle_forward(sd, spec, x)
    assert (got - want).abs().max().item() <= BF16_TOL and O.psnr(got, want, 1.0) >= BF16_PSNR


def test_bf16_conv3_lightweight_matches_reference_vectors_and_screenshot():
    from fs_uae_image_enhancer_project_b200 import model_conv3
    g = load_gold("conv3_lightweight")
    m = model_conv3.get_model("lightweight")
    m.load_state_dict(O.make_bn_state_dict(O.conv3_channels("lightweight"), int(g["seed"])))
    m = m.to(dev()).set_precision("bf16")
    y = m(torch.from_numpy(g["x"]).to(dev())).float().cpu()
    want = torch.from_numpy(g["y"])
    assert y.shape == want.shape and (y[:, 3] == 255.0).all()
    assert (y - want).abs().max().item() <= 255 * BF16_TOL and O.psnr(y, want, 255.0) >= 50.0
    # trained weights on a real screenshot, full frame (6 strips at full resolution)
    m = model_conv3.get_model("lightweight")
    m.load_state_dict(trained_conv3_sd())
    m = m.to(dev()).set_precision("bf16")
    x = load_png_rgba(os.path.join(GOLD, "samples", "sample5.png")).permute(0, 3, 1, 2).contiguous()
    want = load_png_rgba(os.path.join(GOLD, "predicted_conv3", "sample5.png")).permute(0, 3, 1, 2)
    got = m(x.to(dev())).float().cpu().clamp(0, 255).to(torch.uint8)
    assert (got.int() - want.int()).abs().max().item() <= 4 and O.psnr(got, want, 255.0) >= 48.0


@pytest.mark.parametrize("preset", ["lightweight", "heavyweight"])
def test_bf16_conv5_matches_reference_vectors(preset):
    """conv5 heavyweight's 128->128 layers (residual add read from global memory) run on the wide tile kernel."""
    from fs_uae_image_enhancer_project_b200 import model_conv5
    g = load_gold(f"conv5_{preset}")
    m = model_conv5.get_model(preset)
    m.load_state_dict(O.make_bn_state_dict(O.conv5_channels(preset), int(g["seed"])))
    m = m.to(dev()).set_precision("bf16")
    y = m(torch.from_numpy(g["x"]).to(dev())).cpu()
    want = torch.from_numpy(g["y"])
    assert (y - want).abs().max().item() <= BF16_TOL and O.psnr(y, want, 1.0) >= BF16_PSNR


@pytest.mark.parametrize("grid", [1, 2, 3, 5, 7])
def test_bf16_partition_independence(grid, monkeypatch):
    """The persistent CTAs split the strip rows into contiguous ranges; any CTA count must give the
    same bits (regression: a 2-row segment followed by a longer one once read a stale ring row)."""
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 61)
    x = torch.rand(1, 3, 8, 752, generator=torch.Generator().manual_seed(5))
    m = _bf16_model(spec, sd)
    base = m(x.to(dev())).cpu()
    assert (base - O.pix_shuffle_forward(sd, spec, x)).abs().max().item() <= BF16_TOL
    monkeypatch.setenv("FSUAE_DEBUG_GRID", str(grid))
    assert torch.equal(m(x.to(dev())).cpu(), base)


def test_engine_file_and_raw_cli_match_the_module(tmp_path):
    """Deploy path: engine file -> fsuae_engine_create_from_file -> same bytes as the module; raw CLI works."""
    import ctypes as C
    from fs_uae_image_enhancer_project_b200 import _lib, export, raw_framebuffer
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 71)
    m = build_pkg_pix_shuffle(spec, sd).to(dev())
    fb = O.synth_framebuffers(2, seed=3, h=64, w=96)
    want = m.forward_framebuffer(fb.to(dev())).cpu()
    path = str(tmp_path / "m.fsuae")
    export.export_engine_file(m, path)
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.fsuae_engine_create_from_file(path.encode(), 0, _lib.PREC_FP32, 64, 96, 4, C.byref(h)) == _lib.OK
    out = torch.empty_like(fb)
    flags = _lib.FLAG_GAMMA_IN | _lib.FLAG_GAMMA_OUT
    assert lib.fsuae_engine_run_host(h, fb.data_ptr(), out.data_ptr(), 2, _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags) == _lib.OK
    lib.fsuae_engine_destroy(h)
    assert torch.equal(out, want)
    # raw CLI (reference: convert_raw_to_png_using_final_model.py)
    raw = tmp_path / "frames.raw"
    fb.numpy().tofile(raw)
    wpath = tmp_path / "w.pt"
    torch.save(sd, wpath)
    outp = tmp_path / "out.raw"
    assert raw_framebuffer.main([str(wpath), str(raw), str(outp), "--width", "96", "--height", "64", "--precision", "fp32"]) == 0
    got = torch.from_numpy(np.fromfile(outp, dtype=np.uint8).reshape(2, 64, 96, 4))
    assert torch.equal(got, want)


@pytest.mark.parametrize("n", [2, 3, 5])
def test_bf16_cta_pairs_odd_and_even_batches(n):
    """Launches with >= 2 frames use the CTA-pair kernels (two frames in lockstep); an odd count makes the last
    pair compute one frame twice.  Results must equal the single-CTA kernels bit for bit, run after run."""
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 81)
    x = torch.rand(n, 3, 96, 600, generator=torch.Generator().manual_seed(n)).to(dev())
    m = _bf16_model(spec, sd)
    single = torch.cat([m(x[i:i + 1].contiguous()) for i in range(n)])      # n == 1 launches: single-CTA kernels
    for _ in range(5):
        assert torch.equal(m(x), single)
    assert (single.cpu() - O.pix_shuffle_forward(sd, spec, x.cpu())).abs().max().item() <= BF16_TOL


def test_bf16_is_reproducible_at_full_size():
    """Regression for two shared-memory hand-off races found in this build (0*NaN from unlanded ring rows; ring
    rows released before the residual loads had completed): 12 runs of a 4-frame batch, identical bits."""
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 82)
    x = torch.rand(4, 3, 576, 752, generator=torch.Generator().manual_seed(8)).to(dev())
    m = _bf16_model(spec, sd)
    first = m(x).clone()
    for _ in range(11):
        assert torch.equal(m(x), first)


def test_streaming_host_submissions_match_the_blocking_call():
    """fsuae_engine_submit_host x3 + one wait == three fsuae_engine_run_host calls (staging pairs carry across
    submissions; upload, forward and download of consecutive submissions overlap)."""
    from fs_uae_image_enhancer_project_b200 import _lib
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 77)
    m = _bf16_model(spec, sd)
    m.chunk_frames = 8
    fbs = [O.synth_framebuffers(n, seed=20 + n, h=64, w=96).pin_memory() for n in (5, 40, 9)]
    want = [m.run_host(fb) for fb in fbs]
    eng = m.engine_for(dev(), 64, 96)
    outs = [torch.zeros_like(fb).pin_memory() for fb in fbs]
    flags = _lib.FLAG_GAMMA_IN | _lib.FLAG_GAMMA_OUT
    for fb, out in zip(fbs, outs):
        eng.submit_host(fb, out, fb.shape[0], _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags)
    eng.wait_host()
    for w, o in zip(want, outs):
        assert torch.equal(w, o)
    assert (want[0].int() - O.framebuffer_forward(sd, spec, fbs[0]).int()).abs().max().item() <= 6


def test_bf16_three_rows_per_instruction_kernels(monkeypatch):
    """Opt-in R3 kernels (one MMA feeds three output rows, N = 3 x NPAD, circular TMEM accumulators): same results as
    the default CTA-pair kernels within bf16 tolerance, for ragged segment partitions too."""
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 91)
    x = torch.rand(3, 3, 50, 600, generator=torch.Generator().manual_seed(6))
    want = O.pix_shuffle_forward(sd, spec, x)
    monkeypatch.setenv("FSUAE_R3", "1")
    m = _bf16_model(spec, sd)
    got = m(x.to(dev())).cpu()
    assert (got - want).abs().max().item() <= BF16_TOL and O.psnr(got, want, 1.0) >= BF16_PSNR
    for grid in (1, 3, 7):
        monkeypatch.setenv("FSUAE_DEBUG_GRID", str(grid))
        assert torch.equal(m(x.to(dev())).cpu(), got)


def test_bf16_arbitrary_channel_plan_with_skip_projections():
    """Channel plan (24, 40, 40, 56, 20, 28): both short skips go through 1x1 projections (run as centre-tap layers, the
    residual then comes from global memory), no layer has an instantiated resident-weight kernel, so everything --
    including the concat layer and the PixelShuffle tail -- runs on the K-streamed wide kernel."""
    g = load_gold("pix_shuffle_proj_a")
    spec = gold_spec("proj_a")
    sd = O.make_pix_shuffle_state_dict(spec, int(g["seed"]))
    m = _bf16_model(spec, sd)
    got = m(torch.from_numpy(g["x"]).to(dev())).cpu()
    want = torch.from_numpy(g["y"])
    assert (got - want).abs().max().item() <= BF16_TOL and O.psnr(got, want, 1.0) >= BF16_PSNR
    fb = O.synth_framebuffers(3, seed=12, h=70, w=300)
    d = (m.forward_framebuffer(fb.to(dev())).cpu().int() - O.framebuffer_forward(sd, spec, fb).int()).abs()
    assert d.max().item() <= 6 and (d == 0).float().mean().item() >= 0.85


def test_quantize_and_synth_kernels_are_bit_exact():
    """Input-side byte kernels: colour-depth grid + pixel-mode replication against the reference's own outputs
    (tests/golden/quantize.npz) and the oracle; the device-side synthetic stream against its numpy restatement."""
    from fs_uae_image_enhancer_project_b200 import synth
    g = load_gold("quantize")
    img = torch.from_numpy(g["img"])[None].to(dev())
    for cs in ("RGB888", "RGB444", "RGB555", "RGB565", "RGB666"):
        got = synth.quantize_frames(img, cs).cpu().numpy()
        assert np.array_equal(got[0, :, :, :3], g[f"q_{cs}"]) and (got[..., 3] == 255).all()
    for style in ("lores", "lores_laced", "hires", "hires_laced"):
        got = synth.quantize_frames(img, "RGB444", style).cpu().numpy()
        assert np.array_equal(got[0, :, :, :3], g[f"post_{style}"])
    rs = np.random.RandomState(3)
    batch = rs.randint(0, 256, (5, 37, 53, 4)).astype(np.uint8)                    # RGBA input, odd sizes
    for cs, style, e17 in (("RGB444", "lores", True), ("RGB565", "hires", False), ("RGB666", "lores_laced", False)):
        got = synth.quantize_frames(torch.from_numpy(batch).to(dev()), cs, style, expand17=e17).cpu().numpy()
        assert np.array_equal(got, O.quantize_frames(batch, cs, style, expand17=e17))
    assert synth.quantize_frames(torch.zeros((0, 4, 4, 3), dtype=torch.uint8, device=dev())).shape == (0, 4, 4, 4)
    with pytest.raises(ValueError):
        synth.quantize_frames(img, "RGB333")
    for n, h, w, seed, first, e17 in ((9, 576, 752, 1234, 0, True), (5, 33, 47, 2 ** 63 + 11, 6, False)):
        got = synth.synth_rgb444_frames(n, h, w, seed=seed, first_frame=first, expand17=e17, device=dev()).cpu().numpy()
        assert np.array_equal(got, O.synth_rgb444_frames(n, h, w, seed, first, e17))


def test_device_generated_stream_through_the_engine():
    """Frames born on the GPU go straight into the fused forward: same bytes as the oracle gets from the same frames."""
    from fs_uae_image_enhancer_project_b200 import synth
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 13)
    m = build_pkg_pix_shuffle(spec, sd).to(dev())                                  # fp32 build: <= 1 LSB
    fb = synth.synth_rgb444_frames(4, 64, 96, seed=9, device=dev())
    want = O.framebuffer_forward(sd, spec, torch.from_numpy(O.synth_rgb444_frames(4, 64, 96, 9)))
    got = m.forward_framebuffer(fb).cpu()
    assert (got.int() - want.int()).abs().max().item() <= 1


@pytest.mark.parametrize("family", ["conv3", "conv5"])
def test_heavyweight_full_size_builds_agree_and_frames_are_independent(family):
    """BASELINE configs 3-4 at 752x576 (6 strips, 72 eight-row blocks, 2 output-channel groups on the wide kernel): the bf16
    build against the fp32 build of the same engine, and size-independent properties (frame order, chunking)."""
    from fs_uae_image_enhancer_project_b200 import model_conv3, model_conv5
    mod, chans = (model_conv3, O.conv3_channels) if family == "conv3" else (model_conv5, O.conv5_channels)
    sd = O.make_bn_state_dict(chans("heavyweight"), 21)
    g = torch.Generator().manual_seed(8)
    if family == "conv3":
        x = torch.randint(0, 256, (3, 4, 576, 752), dtype=torch.uint8, generator=g).to(dev())
        scale = 255.0
    else:
        x = torch.rand(3, 3, 576, 752, generator=g).to(dev())
        scale = 1.0
    m32 = mod.get_model("heavyweight"); m32.load_state_dict(sd); m32 = m32.to(dev())
    m16 = mod.get_model("heavyweight"); m16.load_state_dict(sd); m16 = m16.to(dev()).set_precision("bf16")
    m16.chunk_frames = 2
    y16 = m16(x).float()
    y32 = m32(x[:1].contiguous()).float()
    assert (y16[:1] - y32).abs().max().item() <= scale * BF16_TOL and O.psnr(y16[:1].cpu(), y32.cpu(), scale) >= 50.0
    assert torch.equal(y16.flip(0), m16(x.flip(0).contiguous()).float())            # frames are independent
    assert torch.equal(y16[2:3], m16(x[2:3].contiguous()).float())                  # chunking is invisible
