"""Skewed class distributions in the team-sorted kernel (cfg2 shape): random points vs a cut at constant y vs a coarse dyadic
grid (low digits all zero).  With static class -> warp ownership the skewed inputs put a whole tile into one class per round."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import itna_b200 as t
npts = 40_000_000
g = t.named_comb_tree((2, 30))
s = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
f = t.rand_itn(s, link_space=16, rng=20262, normalise=True)
out = torch.empty(npts, dtype=torch.float64, device="cuda:0")
xr = torch.rand((npts, 2), dtype=torch.float64, device="cuda:0")
xc = xr.clone(); xc[:, 1] = 0.3141592653589793
xg = torch.floor(xr * 1024) / 1024
for env in ({"TTN_MMA_DEEP": "0"}, {}):
    os.environ.pop("TTN_MMA_DEEP", None)
    os.environ.update(env)
    f.invalidate_plans()
    plan = f.plan()
    for name, x in (("random", xr), ("y constant", xc), ("1024^2 dyadic grid points", xg)):
        best = 1e9
        for _ in range(3):
            o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr())
            best = min(best, o.kernel_ms)
        print(f"{env} {name:28s}: {best:8.3f} ms  {npts / best / 1e6:6.3f} G pts/s", flush=True)
