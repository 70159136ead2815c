import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import enhancer_oracle as O
from tests.util import build_pkg_pix_shuffle
dev = torch.device("cuda", 0)
spec = O.pix_shuffle_preset("lightweight")
sd = O.make_pix_shuffle_state_dict(spec, 31)
H, W, n, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
x = torch.rand(n, 3, H, W, generator=torch.Generator().manual_seed(2)).to(dev)
os.environ["FSUAE_NO_PAIRS"] = "1"
ref = build_pkg_pix_shuffle(spec, sd).to(dev).set_precision("bf16")
ref.chunk_frames = n
truth = ref(x).clone()
del os.environ["FSUAE_NO_PAIRS"]
m = build_pkg_pix_shuffle(spec, sd).to(dev).set_precision("bf16")
m.chunk_frames = n
bad = 0
for i in range(reps):
    o = m(x)
    d = (o - truth).abs()
    if d.max().item() > 0:
        bad += 1
        nz = (d > 0).nonzero()
        print("run", i, "max", round(d.max().item(), 5), "count", nz.shape[0], "frames", torch.unique(nz[:, 0]).tolist(),
              "rows %d-%d cols %d-%d" % (nz[:, 2].min() // 2, nz[:, 2].max() // 2, nz[:, 3].min() // 2, nz[:, 3].max() // 2))
print("bad runs", bad, "of", reps)
