"""Device-resident rates of the chain shapes that run the team-sorted kernel (A/B across builds: LIBTTNEVAL=...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import itna_b200 as t
from itna_b200 import _capi


def run(name, f, ncol, npts):
    plan = f.plan()
    info = plan.info()
    x = torch.rand((npts, ncol), dtype=torch.float64, device="cuda:0")
    out = torch.empty(npts * (2 if info["is_complex"] else 1), dtype=torch.float64, device="cuda:0")
    best = 1e9
    for _ in range(4):
        o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr())
        best = min(best, o.kernel_ms)
    print(f"{name:40s} kernel={_capi.KERNEL_NAMES[o.kernel_used]:8s} {best:9.3f} ms {npts / best / 1e6:8.3f} G points/s", flush=True)


N = 40_000_000
s4 = t.continuous_siteinds(t.named_grid((28, 1)), map_dimension=2)
run("cfg4 shape: 28 sites chi=32 (2 gathers, no rounds)", t.rand_itn(s4, link_space=32, rng=0, normalise=True), 2, N)
s16 = t.continuous_siteinds(t.named_grid((28, 1)), map_dimension=2)
run("28 sites chi=16 (2 gathers, no rounds)", t.rand_itn(s16, link_space=16, rng=0, normalise=True), 2, N)
s41 = t.continuous_siteinds(t.named_grid((41, 1)))
run("41 sites chi=32, 1-D", t.rand_itn(s41, link_space=32, rng=0, normalise=True), 1, N // 2)
g = t.named_comb_tree((2, 30))
s = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
run("cfg2 shape chi=8", t.rand_itn(s, link_space=8, rng=0, normalise=True), 2, N)
for base in (3, 4):
    sb = t.continuous_siteinds(t.named_grid((60, 1)), base=base)
    run(f"60 sites chi=16 base {base}", t.rand_itn(sb, link_space=16, rng=0, normalise=True), 1, N)
