import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import enhancer_oracle as O
from tests.util import build_pkg_pix_shuffle
dev = torch.device("cuda", 0)
spec = O.pix_shuffle_preset("lightweight")
sd = O.make_pix_shuffle_state_dict(spec, 41)
m = build_pkg_pix_shuffle(spec, sd).to(dev).set_precision("bf16")
m.chunk_frames = 4
x = torch.rand(6, 3, 576, 752, generator=torch.Generator().manual_seed(2)).to(dev)
a = m(x[4:5].contiguous()); b = m(x[4:5].contiguous())
print("same call twice equal:", torch.equal(a, b), (a - b).abs().max().item())
y2 = m(x[4:6].contiguous())
d = (y2[0] - a[0]).abs()
print("n=2 vs n=1: max", d.max().item(), "count", (d > 0).sum().item())
nz = (d > 0).nonzero()
if nz.numel():
    print("rows:", nz[:, 1].min().item(), nz[:, 1].max().item(), "cols:", nz[:, 2].min().item(), nz[:, 2].max().item())
    rows = torch.unique(nz[:, 1]); cols = torch.unique(nz[:, 2])
    print("unique rows", rows[:40].tolist(), len(rows)); print("unique cols", cols[:40].tolist(), len(cols))
y4 = m(x[:4].contiguous()); y4b = m(x[:4].contiguous())
print("n=4 twice equal:", torch.equal(y4, y4b), (y4 - y4b).abs().max().item())
want = O.pix_shuffle_forward(sd, spec, x[4:5].cpu())
print("err n=1:", (a.cpu() - want).abs().max().item(), " err n=2:", (y2[:1].cpu() - want).abs().max().item())
y = m(x)
for i in range(6):
    s = m(x[i:i+1].contiguous())
    d = (y[i] - s[0]).abs()
    nz = (d > 0).nonzero()
    print("frame", i, "max diff", d.max().item(), "count", nz.shape[0],
          "rows", (nz[:, 1].min().item(), nz[:, 1].max().item()) if nz.numel() else None,
          "cols", (nz[:, 2].min().item(), nz[:, 2].max().item()) if nz.numel() else None)
    if nz.numel():
        print("   unique rows", torch.unique(nz[:, 1])[:30].tolist())
