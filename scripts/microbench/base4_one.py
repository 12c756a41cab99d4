import os, sys
sys.path.insert(0, os.getcwd())
import torch
import itna_b200 as t
base = int(sys.argv[1]) if len(sys.argv) > 1 else 4
npts = 20_000_000
x = torch.rand((npts, 1), dtype=torch.float64, device="cuda:0")
out = torch.empty(npts, dtype=torch.float64, device="cuda:0")
s = t.continuous_siteinds(t.named_grid((60, 1)), base=base)
f = t.rand_itn(s, link_space=16, rng=base, normalise=True)
plan = f.plan()
for _ in range(3):
    o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr())
print(base, o.kernel_ms)
