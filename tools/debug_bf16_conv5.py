"""Per-layer comparison of the bf16 engine with the fp32 oracle for model_conv5 (debug aid)."""
import sys, os, ctypes, torch, numpy as np
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import enhancer_oracle as O
from fs_uae_image_enhancer_project_b200 import model_conv5
dev = torch.device("cuda", 0)
preset = sys.argv[1]
sd = O.make_bn_state_dict(O.conv5_channels(preset), 42)
m = model_conv5.get_model(preset); m.load_state_dict(sd); m = m.to(dev).set_precision("bf16")
H, W = 36, 52
x = torch.rand(1, 3, H, W, generator=torch.Generator().manual_seed(2))
got = m(x.to(dev)).cpu()
eng = m.engine_for(dev, H, W); lib = eng._lib
lib.fsuae_debug_read_bf16_buffer.restype = ctypes.c_longlong
lib.fsuae_debug_read_bf16_buffer.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong]
bn = lambda i, t: O._bn(sd, i, t, torch.float32)
cv = lambda i, t: F.conv2d(t, sd[f"conv{i}.weight"], None, 1, 1)
l1 = torch.relu(bn(1, cv(1, x))); l2 = torch.relu(l1 + bn(2, cv(2, l1))); l3 = torch.relu(bn(3, cv(3, l2))); l4 = torch.relu(l3 + bn(4, cv(4, l3)))
y = torch.sigmoid(bn(5, cv(5, l4)))
refs = [x, l1, l2, l3, l4]
S = (W + 125) // 126; PW = 126 * S + 4
for i, ref in enumerate(refs):
    C = ref.shape[1]; NP = (C + 7) // 8
    nbytes = NP * (H + 4) * PW * 16
    buf = np.zeros(nbytes, dtype=np.uint8)
    lib.fsuae_debug_read_bf16_buffer(eng._h, i, buf.ctypes.data, nbytes)
    u16 = torch.from_numpy(buf.view(np.uint16).astype(np.int32))
    f32 = (u16 << 16).view(torch.int32).view(torch.float32).view(NP, H + 4, PW, 8)
    mine = f32[:, 2:H + 2, 2:W + 2, :].permute(0, 3, 1, 2).reshape(NP * 8, H, W)[:C]
    e = (mine - ref[0]).abs()
    rel = e.max().item() / ref.abs().max().item()
    worst_c = e.amax(dim=(1, 2)).argmax().item()
    print(f"buffer {i}: C={C} ref max {ref.abs().max().item():.3f} max err {e.max().item():.4f} (rel {rel:.2e}) mean err {e.mean().item():.2e} worst ch {worst_c}")
print("final max err", (got - y).abs().max().item())
