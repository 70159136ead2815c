"""First-run layer-by-layer check of the bf16 engine (all frames) against the fp32 oracle."""
import sys, os, ctypes, torch, numpy as np
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import enhancer_oracle as O
from tests.util import build_pkg_pix_shuffle
dev = torch.device("cuda", 0)
spec = O.pix_shuffle_preset("lightweight")
sd = O.make_pix_shuffle_state_dict(spec, 31)
m = build_pkg_pix_shuffle(spec, sd).to(dev).set_precision("bf16")
H, W, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
m.chunk_frames = n
x = torch.rand(n, 3, H, W, generator=torch.Generator().manual_seed(2))
got = m(x.to(dev)).cpu()
eng = m.engine_for(dev, H, W); lib = eng._lib
lib.fsuae_debug_read_bf16_buffer.restype = ctypes.c_longlong
lib.fsuae_debug_read_bf16_buffer.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong]
act = lambda slot, t: O.apply_activation(spec.acts[slot][0], t, sd, slot, spec.acts[slot][1])
conv = lambda i, t: F.conv2d(t, sd[f"conv{i}.weight"], sd[f"conv{i}.bias"], padding=1)
t0 = F.pixel_unshuffle(x, 2)
l1 = act("l1_act2", act("l1_act1", conv(1, t0)))
l2 = act("l2_act4", act("l2_act3", l1 + act("l2_act2", act("l2_act1", conv(2, l1)))))
l3 = conv(3, l2)
l4 = act("l4_act4", act("l4_act3", l3 + act("l4_act2", act("l4_act1", conv(4, l3)))))
l5 = conv(5, l4)
l6 = act("l6_act2", act("l6_act1", conv(6, torch.cat([l1, l5], 1))))
refs = [t0, l1, l2, l3, l4, l5, l6]
Hw, Ww = H // 2, W // 2
S = (Ww + 125) // 126; PW = 126 * S + 4
for i, ref in enumerate(refs):
    C = ref.shape[1]; NP = (C + 7) // 8 if i else 2
    fbytes = NP * (Hw + 4) * PW * 16
    buf = np.zeros(fbytes * n, dtype=np.uint8)
    lib.fsuae_debug_read_bf16_buffer(eng._h, i, buf.ctypes.data, fbytes * n)
    u16 = torch.from_numpy(buf.view(np.uint16).astype(np.int32))
    f32 = (u16 << 16).view(torch.int32).view(torch.float32).view(n, NP, Hw + 4, PW, 8)
    mine = f32[:, :, 2:Hw + 2, 2:Ww + 2, :].permute(0, 1, 4, 2, 3).reshape(n, NP * 8, Hw, Ww)[:, :C]
    e = (mine - ref).abs().amax(dim=1)
    thr = 0.03 * max(1.0, ref.abs().max().item())
    bad = (e > thr).nonzero()
    msg = ""
    if bad.numel():
        msg = "frames %s rows %d-%d cols %d-%d" % (torch.unique(bad[:, 0]).tolist(), bad[:, 1].min(), bad[:, 1].max(), bad[:, 2].min(), bad[:, 2].max())
    print(f"buffer {i}: max err {e.max().item():.4f} bad px {bad.shape[0]} {msg}")
print("final err", (got - O.pix_shuffle_forward(sd, spec, x)).abs().max().item())
