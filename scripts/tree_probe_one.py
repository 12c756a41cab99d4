"""One tree workload for a launch list (ncu): real 3-tooth comb of 60 vertices, chi = 16, 2e6 points."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itna_b200 as t
L = 20
g = t.named_comb_tree((3, L))
s = t.continuous_siteinds(g, [[(j, i) for i in range(1, L + 1)] for j in range(1, 4)])
f = t.rand_itn(s, link_space=16, rng=16, normalise=True)
plan = f.plan()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
x = torch.rand((n, 3), dtype=torch.float64, device="cuda:0")
out = torch.empty(n, dtype=torch.float64, device="cuda:0")
for _ in range(2):
    o = plan.evaluate_device(x.data_ptr(), n, out.data_ptr())
print(o.kernel_ms, "ms", o.n_launches, "launches")
