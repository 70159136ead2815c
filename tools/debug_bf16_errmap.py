import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import enhancer_oracle as O
from tests.util import build_pkg_pix_shuffle
dev = torch.device("cuda", 0)
spec = O.pix_shuffle_preset("lightweight")
sd = O.make_pix_shuffle_state_dict(spec, 41)
m = build_pkg_pix_shuffle(spec, sd).to(dev).set_precision("bf16")
H, W = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1
x = torch.rand(n, 3, H, W, generator=torch.Generator().manual_seed(2))
want = O.pix_shuffle_forward(sd, spec, x)
got = m(x.to(dev)).cpu()
e = (got - want).abs().amax(dim=1)            # [n,H,W]
print("max err", e.max().item())
for f in range(n):
    bad = (e[f] > 3e-3).nonzero()
    if bad.numel() == 0:
        print("frame", f, "clean"); continue
    rows = torch.unique(bad[:, 0] // 2).tolist(); cols = torch.unique(bad[:, 1] // 2).tolist()
    def runs(v):
        out = []; s = v[0]; p = v[0]
        for a in v[1:]:
            if a != p + 1: out.append((s, p)); s = a
            p = a
        out.append((s, p)); return out
    print("frame", f, "bad half-res rows", runs(rows), "cols", runs(cols))
