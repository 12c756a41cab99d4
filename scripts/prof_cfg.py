"""One evaluation of BASELINE config 3 or 5 (for ncu): python scripts/prof_cfg.py 3|5 npts"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itna_b200 as t
cfg, npts = int(sys.argv[1]), int(float(sys.argv[2]))
if cfg == 3:
    g = t.named_binary_tree(7)
    ws = g.vertices()[7:]
    s = t.continuous_siteinds(g, [ws[i::3] for i in range(3)])
    f, ncol, nout = t.rand_itn(s, link_space=64, rng=3, normalise=True), 3, 1
else:
    s = t.complex_continuous_siteinds(t.named_grid((40, 1)), map_dimension=2)
    f, ncol, nout = t.rand_itn(s, link_space=128, rng=5, eltype=complex, normalise=True), 4, 2
plan = f.plan()
x = torch.rand((npts, ncol), dtype=torch.float64, device="cuda:0")
out = torch.empty(nout * npts, dtype=torch.float64, device="cuda:0")
for _ in range(2):
    o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr())
print(cfg, npts, o.kernel_ms, "ms", o.n_launches, "launches")
