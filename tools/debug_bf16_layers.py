"""Compare every intermediate activation map of the bf16 engine with the fp32 oracle (debug aid)."""
import sys, os, ctypes, torch, numpy as np
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import enhancer_oracle as O
from tests.util import build_pkg_pix_shuffle
dev = torch.device("cuda", 0)
spec = O.pix_shuffle_preset("lightweight")
sd = O.make_pix_shuffle_state_dict(spec, 41)
m = build_pkg_pix_shuffle(spec, sd).to(dev).set_precision("bf16")
H, W = int(sys.argv[1]), int(sys.argv[2])
x = torch.rand(1, 3, H, W, generator=torch.Generator().manual_seed(2))
got = m(x.to(dev)).cpu()
eng = m.engine_for(dev, H, W)
lib = eng._lib
lib.fsuae_debug_read_bf16_buffer.restype = ctypes.c_longlong
lib.fsuae_debug_read_bf16_buffer.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong]
# oracle intermediates
act = lambda slot, t: O.apply_activation(spec.acts[slot][0], t, sd, slot, spec.acts[slot][1])
conv = lambda i, t: F.conv2d(t, sd[f"conv{i}.weight"], sd[f"conv{i}.bias"], padding=1)
t0 = F.pixel_unshuffle(x, 2)
l1 = act("l1_act2", act("l1_act1", conv(1, t0)))
l2 = act("l2_act4", act("l2_act3", l1 + act("l2_act2", act("l2_act1", conv(2, l1)))))
l3 = act("l3_act2", act("l3_act1", conv(3, l2)))
l4 = act("l4_act4", act("l4_act3", l3 + act("l4_act2", act("l4_act1", conv(4, l3)))))
l5 = conv(5, l4)
l6 = act("l6_act2", act("l6_act1", conv(6, torch.cat([l1, l5], 1))))
refs = [t0, l1, l2, l3, l4, l5, l6]
Hw, Ww = H // 2, W // 2
S = (Ww + 125) // 126
PW = 126 * S + 4
for i, ref in enumerate(refs):
    C = ref.shape[1]
    NP = (C + 7) // 8 if i else 2
    nbytes = NP * (Hw + 4) * PW * 16
    buf = np.zeros(nbytes, dtype=np.uint8)
    n = lib.fsuae_debug_read_bf16_buffer(eng._h, i, buf.ctypes.data, nbytes)
    a = torch.from_numpy(buf.view(np.int16).astype(np.int32) << 16).view(torch.float32) if False else None
    u16 = torch.from_numpy(buf.view(np.uint16).astype(np.int32))
    f32 = (u16 << 16).view(torch.int32).view(torch.float32).view(NP, Hw + 4, PW, 8)
    mine = f32[:, 2:Hw + 2, 2:Ww + 2, :].permute(0, 3, 1, 2).reshape(NP * 8, Hw, Ww)[:C]
    e = (mine - ref[0]).abs().amax(dim=0)
    bad = (e > 0.02 * max(1.0, ref.abs().max().item())).nonzero()
    border_ok = (f32[:, 0].abs().max().item() == 0 and f32[:, Hw + 3].abs().max().item() == 0 and f32[:, :, 0].abs().max().item() == 0)
    print(f"buffer {i}: C={C} max err {e.max().item():.4f} (ref max {ref.abs().max().item():.2f}) bad px {bad.shape[0]} border_zero={border_ok}",
          ("rows %d-%d cols %d-%d" % (bad[:, 0].min(), bad[:, 0].max(), bad[:, 1].min(), bad[:, 1].max())) if bad.numel() else "")
