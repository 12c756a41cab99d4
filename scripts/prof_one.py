"""Run ONE workload a few times (for ncu): python scripts/prof_one.py cfg2 dmma 2e6"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itna_b200 as t

cfg, kernel, npts = sys.argv[1], sys.argv[2], int(float(sys.argv[3]))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
if cfg == "cfg2":
    g = t.named_comb_tree((2, 30))
    s = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
    f, ncol = t.rand_itn(s, link_space=16, rng=0, normalise=True), 2
elif cfg == "cfg4":
    s = t.continuous_siteinds(t.named_grid((28, 1)), map_dimension=2)
    f, ncol = t.rand_itn(s, link_space=32, rng=0, normalise=True), 2
plan = f.plan()
x = torch.rand((npts, ncol), dtype=torch.float64, device="cuda:0")
out = torch.empty(npts, dtype=torch.float64, device="cuda:0")
for _ in range(reps):
    o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr(), kernel=kernel)
print(cfg, kernel, npts, o.kernel_ms, "ms")
