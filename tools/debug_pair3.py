"""Loop pair-mode forwards until the output differs from the single-CTA result, then locate the first bad layer."""
import sys, os, ctypes, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import enhancer_oracle as O
from tests.util import build_pkg_pix_shuffle
dev = torch.device("cuda", 0)
spec = O.pix_shuffle_preset("lightweight")
sd = O.make_pix_shuffle_state_dict(spec, 31)
H, W, n = 576, 752, 4
x = torch.rand(n, 3, H, W, generator=torch.Generator().manual_seed(2)).to(dev)
Hw, Ww = H // 2, W // 2
S = (Ww + 125) // 126; PW = 126 * S + 4
planes = [2, 5, 5, 9, 9, 5, 5]

def read_all(m):
    eng = m.engine_for(dev, H, W); lib = eng._lib
    lib.fsuae_debug_read_bf16_buffer.restype = ctypes.c_longlong
    lib.fsuae_debug_read_bf16_buffer.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong]
    out = []
    for i, NP in enumerate(planes):
        nb = NP * (Hw + 4) * PW * 16 * n
        buf = np.zeros(nb, dtype=np.uint8)
        lib.fsuae_debug_read_bf16_buffer(eng._h, i, buf.ctypes.data, nb)
        out.append(torch.from_numpy(buf.view(np.int16).copy()).view(n, NP, Hw + 4, PW, 8))
    return out

os.environ["FSUAE_NO_PAIRS"] = "1"
ref = build_pkg_pix_shuffle(spec, sd).to(dev).set_precision("bf16"); ref.chunk_frames = n
truth = ref(x).clone(); tb = read_all(ref)
del os.environ["FSUAE_NO_PAIRS"]
m = build_pkg_pix_shuffle(spec, sd).to(dev).set_precision("bf16"); m.chunk_frames = n
for it in range(40):
    o = m(x)
    if (o - truth).abs().max().item() > 0:
        mb = read_all(m)
        for i in range(7):
            d = (mb[i] != tb[i]).nonzero()
            if d.numel():
                fr = torch.unique(d[:, 0]).tolist(); pl = torch.unique(d[:, 1]).tolist()
                print(f"iter {it} buffer {i}: {d.shape[0]} differing values frames {fr} planes {pl} rows {d[:,2].min()-1}-{d[:,2].max()-1} cols {d[:,3].min()-1}-{d[:,3].max()-1} ch-in-chunk {torch.unique(d[:,4]).tolist()}")
                # detail for the first bad buffer
                rows = torch.unique(d[:, 2]).tolist(); cols = torch.unique(d[:, 3]).tolist()
                print("   rows", [r - 1 for r in rows][:12], "cols", [c - 1 for c in cols][:40])
                break
