"""CPU: host-side logic -- the C-ABI library loads and exports what include/fsuae_enhancer.h
declares, descriptors / state_dict keys / error behaviour mirror the reference."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import enhancer_oracle as O
from tests.util import build_pkg_pix_shuffle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from fs_uae_image_enhancer_project_b200 import _lib
    header = open(os.path.join(ROOT, "include", "fsuae_enhancer.h")).read()
    declared = set(re.findall(r"FSUAE_API\s+[\w\s\*]+?\b(fsuae_\w+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), (declared, set(_lib.EXPORTS))
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.fsuae_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header():
    from fs_uae_image_enhancer_project_b200 import _lib
    assert ctypes.sizeof(_lib.ActDesc) == 20
    assert ctypes.sizeof(_lib.LayerDesc) == 40 + 8 * 20 + 8
    assert ctypes.sizeof(_lib.NetDesc) == 24 + 16 * ctypes.sizeof(_lib.LayerDesc)


def test_create_without_gpu_fails_loudly():
    """No CPU fallback: on a box without a GPU the engine refuses to exist."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fs_uae_image_enhancer_project_b200 import _lib, model_pix_shuffle
    from fs_uae_image_enhancer_project_b200.descriptor import build_descriptor
    from fs_uae_image_enhancer_project_b200.engine import Engine
    m = model_pix_shuffle.get_model("lightweight")
    desc, blob = build_descriptor(m._layer_specs(), m._head, m._tail)
    with pytest.raises(_lib.EngineError) as ei:
        Engine(desc, blob, 0, _lib.PREC_FP32, 576, 752)
    assert ei.value.code == _lib.ERR_NO_DEVICE
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 8, 8))


def test_invalid_descriptor_is_rejected_before_touching_the_gpu():
    from fs_uae_image_enhancer_project_b200 import _lib, model_pix_shuffle
    from fs_uae_image_enhancer_project_b200.descriptor import build_descriptor
    from fs_uae_image_enhancer_project_b200.engine import Engine
    m = model_pix_shuffle.get_model("lightweight")
    desc, blob = build_descriptor(m._layer_specs(), m._head, m._tail)
    with pytest.raises(ValueError, match="even"):
        Engine(desc, blob, 0, _lib.PREC_FP32, 575, 752)
    desc.layers[2].cin0 = 35
    with pytest.raises(ValueError, match="src0"):
        Engine(desc, blob, 0, _lib.PREC_FP32, 576, 752)


@pytest.mark.parametrize("preset", ["lightweight", "heavyweight"])
def test_pix_shuffle_state_dict_keys_match_reference(preset):
    from fs_uae_image_enhancer_project_b200 import model_pix_shuffle
    m = model_pix_shuffle.get_model(preset)
    spec = O.pix_shuffle_preset(preset)
    sd = O.make_pix_shuffle_state_dict(spec, 1)       # key set == reference state_dict (checked in gen_golden)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    assert sum(p.numel() for p in m.parameters()) == {"lightweight": 136602, "heavyweight": 218105}[preset]
    assert model_pix_shuffle.get_model("nope") is None


@pytest.mark.parametrize("family", ["pix_shuffle", "conv3", "conv5"])
def test_genuine_reference_checkpoints_load_with_strict_true(family):
    """train.py:236/246 saves model.state_dict() of a Model that owns `perceptual_criterion` (a PerceptualLoss with a
    torchvision VGG16, loss_vgg.py:60): real best.pth files carry those keys.  README quick-start:
    m.load_state_dict(torch.load('best.pth')) -- default strict=True -- must work."""
    import json
    from fs_uae_image_enhancer_project_b200 import model_conv3, model_conv5, model_pix_shuffle
    from tests.util import GOLD
    extra = json.load(open(os.path.join(GOLD, "reference_checkpoint_loss_keys.json")))
    if family == "pix_shuffle":
        m = model_pix_shuffle.get_model("lightweight")
        sd = O.make_pix_shuffle_state_dict(O.pix_shuffle_preset("lightweight"), 3)
    else:
        mod, chans = (model_conv3, O.conv3_channels) if family == "conv3" else (model_conv5, O.conv5_channels)
        m = mod.get_model("lightweight")
        sd = O.make_bn_state_dict(chans("lightweight"), 3)
    ckpt = dict(sd)
    for k, shape in zip(extra["keys"], extra["shapes"]):
        ckpt[k] = torch.zeros([min(d, 2) for d in shape])           # contents are irrelevant: the entries are dropped
    assert any(k.startswith("perceptual_criterion.vgg.features.") for k in ckpt)
    res = m.load_state_dict(ckpt)                                   # strict=True
    assert not res.missing_keys and not res.unexpected_keys
    assert all(torch.equal(m.state_dict()[k], v) for k, v in sd.items())
    with pytest.raises(RuntimeError, match="Unexpected key"):
        m.load_state_dict({**sd, "bogus.weight": torch.zeros(1)})   # anything else unexpected still raises


def test_parameter_dtype_selects_the_build():
    from fs_uae_image_enhancer_project_b200 import _lib, model_pix_shuffle
    m = model_pix_shuffle.get_model("lightweight")
    assert m._precision() == _lib.PREC_FP32
    assert m.half()._precision() == _lib.PREC_FP16          # what the reference deploys (torch2onnx.py:58)
    assert m.bfloat16()._precision() == _lib.PREC_BF16
    with pytest.raises(TypeError, match="fp32, fp16 and bf16"):
        m.double()._precision()
    assert m.set_precision("fp16")._precision() == _lib.PREC_FP16
    with pytest.raises(KeyError):
        m.set_precision("int8")


def test_create_rejects_bad_precision_and_oversized_chunk():
    from fs_uae_image_enhancer_project_b200 import _lib, model_pix_shuffle
    from fs_uae_image_enhancer_project_b200.descriptor import build_descriptor
    from fs_uae_image_enhancer_project_b200.engine import Engine
    m = model_pix_shuffle.get_model("lightweight")
    desc, blob = build_descriptor(m._layer_specs(), m._head, m._tail)
    with pytest.raises(ValueError, match="precision"):
        Engine(desc, blob, 0, 7, 576, 752)
    with pytest.raises(ValueError, match="max_chunk_frames"):
        Engine(desc, blob, 0, _lib.PREC_FP32, 576, 752, _lib.MAX_CHUNK_FRAMES + 1)


def test_conv3_conv5_state_dict_keys_and_param_counts():
    from fs_uae_image_enhancer_project_b200 import model_conv3, model_conv5
    for mod, chans, counts in ((model_conv3, O.conv3_channels, (21222, 455366)),
                               (model_conv5, O.conv5_channels, (67494, 264006))):
        for preset, want in zip(("lightweight", "heavyweight"), counts):
            m = mod.get_model(preset)
            sd = O.make_bn_state_dict(chans(preset), 1)
            assert set(m.state_dict()) == set(sd)
            assert sum(p.numel() for p in m.parameters()) == want   # incl. BN affine


def test_descriptor_contents_lightweight():
    from fs_uae_image_enhancer_project_b200 import _lib
    from fs_uae_image_enhancer_project_b200.descriptor import build_descriptor
    spec = O.pix_shuffle_preset("lightweight")
    sd = O.make_pix_shuffle_state_dict(spec, 4)
    m = build_pkg_pix_shuffle(spec, sd)
    d, blob = build_descriptor(m._layer_specs(), m._head, m._tail)
    assert d.n_layers == 7 and d.head == _lib.HEAD_UNSHUFFLE2 and d.tail == _lib.TAIL_SHUFFLE2_RESIDUAL_RELU
    L = d.layers
    assert [(L[i].cin0, L[i].cin1, L[i].cout) for i in range(7)] == \
        [(12, 0, 36), (36, 0, 36), (36, 0, 72), (72, 0, 72), (72, 0, 36), (36, 36, 36), (36, 0, 12)]
    assert [L[i].skip_src for i in range(7)] == [-1, 1, -1, 3, -1, -1, -1]
    assert (L[5].src0, L[5].src1) == (1, 5)
    # identity slots are dropped; l2: telu | sinlu, biased_prelu ; l4: mish, biased_prelu | tanh, relu
    assert [L[1].pre[0].op, L[1].post[0].op, L[1].post[1].op] == [_lib.ACT["telu"], _lib.ACT["sinlu"], _lib.ACT["biased_prelu"]]
    assert (L[1].n_pre, L[1].n_post, L[3].n_pre, L[3].n_post, L[2].n_pre, L[6].n_pre) == (1, 2, 2, 2, 0, 1)
    w4 = blob[L[3].w_off:L[3].w_off + 72 * 72 * 9].reshape(72, 72, 3, 3)
    assert np.array_equal(w4, sd["conv4.weight"].numpy())
    a = L[3].pre[1]
    assert a.n0 == 72 and a.n1 == 72
    assert np.array_equal(blob[a.p0_off:a.p0_off + 72], sd["l4_act2.bias"].numpy())
    assert np.array_equal(blob[a.p1_off:a.p1_off + 72], sd["l4_act2.prelu.weight"].numpy())
    assert L[6].pre[0].n0 == 1 and blob[L[6].pre[0].p0_off] == sd["l7_act2.bias"].item()


def test_batchnorm_folding_matches_eval_batchnorm():
    from fs_uae_image_enhancer_project_b200 import model_conv3
    from fs_uae_image_enhancer_project_b200.descriptor import build_descriptor
    m = model_conv3.get_model("lightweight")
    sd = O.make_bn_state_dict(O.conv3_channels("lightweight"), 9)
    m.load_state_dict(sd)
    d, blob = build_descriptor(m._layer_specs(), m._head, m._tail)
    x = torch.rand(1, 3, 12, 12)
    w = torch.from_numpy(blob[d.layers[0].w_off:d.layers[0].w_off + 32 * 27].reshape(32, 3, 3, 3))
    b = torch.from_numpy(blob[d.layers[0].b_off:d.layers[0].b_off + 32])
    folded = torch.nn.functional.conv2d(x, w, b, padding=1)
    want = O._bn(sd, 1, torch.nn.functional.conv2d(x, sd["conv1.weight"], None, padding=1), torch.float32)
    assert (folded - want).abs().max() < 1e-5


def test_activation_factory_error_behaviour():
    from fs_uae_image_enhancer_project_b200 import activations
    with pytest.raises(ValueError, match="Unsupported activation"):
        activations.get_activation("nonsense")
    with pytest.raises(TypeError):
        activations.get_activation("relu", params={"negative_slope": 0.1})
    assert activations.get_activation("SoftMax").dim == 1
    assert activations.get_activation("swish").op_name == "silu"
    assert set(activations.ACTIVATION_REGISTRY) == set(O.ACTIVATION_NAMES)
    bp = activations.get_activation("biased_prelu", {"num_parameters": 5})
    assert set(dict(bp.named_parameters())) == {"bias", "prelu.weight"}
    assert float(bp.bias.abs().max()) <= 0.1 and float(bp.prelu.weight[0]) == 0.25
    with pytest.raises(RuntimeError):
        bp(torch.zeros(1, 5, 2, 2))


def test_model_constructor_errors():
    from fs_uae_image_enhancer_project_b200 import model_conv3, model_pix_shuffle
    with pytest.raises(ValueError, match="odd"):
        model_pix_shuffle.Model(layer3_kernel_size=4)
    with pytest.raises(ValueError, match="odd"):
        model_conv3.Model(kernel_size=2)
    with pytest.raises(ValueError, match="1, 3, 5 and 7"):
        model_pix_shuffle.Model(layer2_kernel_size=9)
    m5 = model_pix_shuffle.Model(layer2_kernel_size=5, layer5_kernel_size=1, layer7_kernel_size=7)     # reference :21-64, 108-115
    assert tuple(m5.conv2.weight.shape) == (36, 36, 5, 5) and m5.conv2.padding == (2, 2) and tuple(m5.conv5.weight.shape) == (36, 36, 1, 1)
    from fs_uae_image_enhancer_project_b200.descriptor import build_descriptor
    d5, blob5 = build_descriptor(m5._layer_specs(), m5._head, m5._tail)
    assert [d5.layers[i].ksize for i in range(7)] == [3, 5, 3, 3, 3, 3, 7]                             # 1x1 = centre tap of 3x3
    w5 = blob5[d5.layers[4].w_off:d5.layers[4].w_off + 36 * 36 * 9].reshape(36, 36, 3, 3)
    assert np.array_equal(w5[:, :, 1, 1], m5.conv5.weight.detach().numpy()[:, :, 0, 0]) and np.count_nonzero(w5) == np.count_nonzero(w5[:, :, 1, 1])
    # channel plans with 1x1 skip projections: same parameter names as the reference (:126-128, :143-145), and the
    # projection becomes a layer of its own in the engine descriptor
    m = model_pix_shuffle.Model(layer1_out_channels=24, layer2_out_channels=36, layer3_out_channels=40, layer4_out_channels=40)
    assert tuple(m.state_dict()["skip1_proj_conv.weight"].shape) == (36, 24, 1, 1) and m.skip2_proj_conv is None
    specs = m._layer_specs()
    assert len(specs) == 8 and specs[1].bias is None and specs[2].skip_src == 2 and specs[2].src0 == 1
    assert float(specs[1].weight[:, :, 0, 0].abs().sum()) == 0.0 and specs[1].weight.shape == (36, 24, 3, 3)
    with pytest.raises(ValueError, match="Unsupported activation"):
        model_pix_shuffle.Model(layer1_act1="nonsense")


def test_residual_feature_block_keys_and_descriptor():
    """Drop-in for residual_feature_block.py:5-55: the reference's parameter names, and the block as descriptor layers."""
    from fs_uae_image_enhancer_project_b200 import _lib, residual_feature_block
    from fs_uae_image_enhancer_project_b200.descriptor import build_descriptor
    from tests.util import build_pkg_residual_block
    m, sd, acts, g = build_pkg_residual_block("resblock_b")                  # 20 -> 32 -> 40 channels, 5x5, projection
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    d, blob = build_descriptor(m._layer_specs(), m._head, m._tail, m._in_channels)
    assert (d.head, d.tail, d.in_channels, d.n_layers) == (_lib.HEAD_FEATURES, _lib.TAIL_FEATURES, 20, 4)
    L = d.layers
    assert [(L[i].cin0, L[i].cout, L[i].ksize, L[i].src0, L[i].skip_src) for i in range(4)] == \
        [(20, 32, 3, 0, -1), (32, 32, 5, 1, -1), (20, 40, 3, 0, -1), (32, 40, 3, 2, 3)]
    assert (L[1].n_pre, L[3].n_pre, L[3].n_post) == (2, 1, 1)
    m2, _, _, _ = build_pkg_residual_block("resblock_a")                      # in == out: identity skip, no projection
    assert m2.proj_conv is None and [s.skip_src for s in m2._layer_specs()] == [-1, -1, 0]
    with pytest.raises(ValueError, match="odd"):
        residual_feature_block.ResidualFeatureBlock(8, 8, 8, 4)
    blk = residual_feature_block.ResidualFeatureBlock(8, 4, 8, 3, acts={"act1": "prelu", "act1_params": {"num_parameters": "channel"},
                                                                         "act2": "identity", "act2_params": None,
                                                                         "act3": "prelu", "act3_params": {"num_parameters": "global"},
                                                                         "act4": "relu", "act4_params": None})
    assert blk.act1.weight.numel() == 4 and blk.act3.weight.numel() == 1      # reference :24-35


def test_onnx_wire_reader_roundtrip(tmp_path):
    """The hand-rolled protobuf reader on a hand-assembled ModelProto{graph{initializer, node}}."""
    from fs_uae_image_enhancer_project_b200 import onnx_weights as ow

    def varint(n):
        out = b""
        while True:
            b = n & 0x7F
            n >>= 7
            out += bytes([b | (0x80 if n else 0)])
            if not n:
                return out

    def ld(fno, payload):
        return varint((fno << 3) | 2) + varint(len(payload)) + payload

    arr = np.arange(6, dtype=np.float16).reshape(2, 3)
    tensor = varint(1 << 3) + varint(2) + varint(1 << 3) + varint(3) + varint(2 << 3) + varint(10) \
        + ld(8, b"conv1.bias") + ld(9, arr.tobytes())
    node = ld(1, b"x") + ld(1, b"conv1.bias") + ld(2, b"y") + ld(4, b"PRelu")
    model = ld(7, ld(1, node) + ld(5, tensor))
    p = tmp_path / "m.onnx"
    p.write_bytes(model)
    inits, nodes = ow.read_onnx_initializers(str(p))
    assert np.array_equal(inits["conv1.bias"], arr) and nodes == [("PRelu", ["x", "conv1.bias"], ["y"])]


def test_engine_file_roundtrip_and_create_from_file_errors(tmp_path):
    from fs_uae_image_enhancer_project_b200 import _lib, export, model_pix_shuffle
    import ctypes as C
    m = model_pix_shuffle.get_model("lightweight")
    path = str(tmp_path / "light.fsuae")
    n = export.export_engine_file(m, path)
    assert n == os.path.getsize(path)
    desc, blob = export.read_engine_file(path)
    assert desc.n_layers == 7 and blob.size >= 136602
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.fsuae_engine_create_from_file(path.encode(), 0, _lib.PREC_FP32, 576, 752, 4, C.byref(h))
    if not torch.cuda.is_available():
        assert rc == _lib.ERR_NO_DEVICE                      # parsed fine, then refused: no CPU path
    bad = tmp_path / "bad.fsuae"
    bad.write_bytes(b"NOTANENGINE" * 10)
    assert lib.fsuae_engine_create_from_file(str(bad).encode(), 0, _lib.PREC_FP32, 576, 752, 4, C.byref(h)) == _lib.ERR_INVALID
    assert b"malformed" in lib.fsuae_last_error(None)
    with pytest.raises(ValueError):
        export.read_engine_file(str(bad))
    # a header whose float count disagrees with the file (truncated blob / absurd count) is refused before any allocation
    good = open(path, "rb").read()
    for name, data in (("trunc.fsuae", good[:-4000]),
                       ("huge.fsuae", good[:12] + (0xFFFFFFF0).to_bytes(4, "little") + good[16:])):
        q = tmp_path / name
        q.write_bytes(data)
        assert lib.fsuae_engine_create_from_file(str(q).encode(), 0, _lib.PREC_FP32, 576, 752, 4, C.byref(h)) == _lib.ERR_INVALID
        assert b"malformed" in lib.fsuae_last_error(None)


def test_raw_framebuffer_loader_validates_size(tmp_path):
    from fs_uae_image_enhancer_project_b200 import raw_framebuffer
    p = tmp_path / "f.raw"
    np.zeros(752 * 576 * 4 * 2, np.uint8).tofile(p)
    assert raw_framebuffer.load_raw_rgba(str(p)).shape == (2, 576, 752, 4)
    np.zeros(1000, np.uint8).tofile(p)
    with pytest.raises(ValueError, match="Expected raw file"):
        raw_framebuffer.load_raw_rgba(str(p))
    assert raw_framebuffer.main(["missing.pt", str(p)]) == 1      # reference tool: print + exit 1
