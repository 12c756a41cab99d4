"""Base-b chains through the planner's kernel: chi = 16 MPS of 60 / 40 sites and chi = 2 / 1 chains,
base 2 vs 3 vs 4, device-resident points/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itna_b200 as t
from itna_b200 import _capi
npts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 40_000_000
x = torch.rand((npts, 1), dtype=torch.float64, device="cuda:0")
out = torch.empty(npts, dtype=torch.float64, device="cuda:0")
for n, chi in ((60, 16), (40, 16), (40, 2), (40, 1)):
    for base in (2, 3, 4):
        s = t.continuous_siteinds(t.named_grid((n, 1)), base=base)
        f = t.rand_itn(s, link_space=chi, rng=base, normalise=True) if chi > 1 else t.exp_itn(s, k=0.7, a=0.1, c=1.0, dim=1)
        plan = f.plan()
        best = 1e9
        for _ in range(4):
            o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr())
            best = min(best, o.kernel_ms)
        print(f"{n} sites chi={chi:2d} base {base}: kernel={_capi.KERNEL_NAMES[o.kernel_used]:6s} {best:8.3f} ms {npts / best / 1e6:8.3f} G pts/s "
              f"executed {o.flops_executed / npts:7.0f} flop/pt", flush=True)
