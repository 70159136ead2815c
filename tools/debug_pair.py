import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import enhancer_oracle as O
from tests.util import build_pkg_pix_shuffle
dev = torch.device("cuda", 0)
spec = O.pix_shuffle_preset("lightweight")
sd = O.make_pix_shuffle_state_dict(spec, 31)
m = build_pkg_pix_shuffle(spec, sd).to(dev).set_precision("bf16")
H, W, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = torch.rand(n, 3, H, W, generator=torch.Generator().manual_seed(2)).to(dev)
outs = [m(x).clone() for _ in range(4)]
want = O.pix_shuffle_forward(sd, spec, x.cpu())
print("err vs oracle", (outs[0].cpu() - want).abs().max().item())
for i in range(1, 4):
    d = (outs[i] - outs[0]).abs()
    nz = (d > 0).nonzero()
    print("run", i, "max diff", d.max().item(), "count", nz.shape[0],
          ("frames %s rows %d-%d cols %d-%d" % (torch.unique(nz[:, 0]).tolist(), nz[:, 2].min(), nz[:, 2].max(), nz[:, 3].min(), nz[:, 3].max())) if nz.numel() else "")
os.environ["FSUAE_NO_PAIRS"] = "1"
m2 = build_pkg_pix_shuffle(spec, sd).to(dev).set_precision("bf16")
o2 = m2(x)
print("pair vs single-CTA kernels: max diff", (o2 - outs[0]).abs().max().item())
