"""Run one small-chi workload a few times (for ncu): python scripts/prof_table.py chi1|chi2|chi4|sin [npts]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itna_b200 as t

which = sys.argv[1]
npts = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
g = t.named_comb_tree((2, 30))
s2 = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
if which == "chi1":
    f, ncol = t.exp_itn(s2, k=0.9, a=0.1, c=1.2, dim=1), 2
elif which == "chi2":
    f, ncol = t.rand_itn(s2, link_space=2, rng=20267, normalise=True), 2
elif which == "chi4":
    f, ncol = t.rand_itn(s2, link_space=4, rng=20268, normalise=True), 2
else:
    f, ncol = t.sin_itn(t.continuous_siteinds(t.named_grid((20, 1))), k=2.0, a=0.3, c=1.1), 1
plan = f.plan()
nc = 2 if plan.info()["is_complex"] else 1
x = torch.rand((npts, ncol), dtype=torch.float64, device="cuda:0")
out = torch.empty(npts * nc, dtype=torch.float64, device="cuda:0")
for _ in range(4):
    o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr(), kernel="table")
print(which, npts, o.kernel_ms, "ms")
