import os, sys
sys.path.insert(0, os.getcwd())
import torch
import itna_b200 as t
npts = 40_000_000
x = torch.rand((npts, 1), dtype=torch.float64, device="cuda:0")
out = torch.empty(npts, dtype=torch.float64, device="cuda:0")
for base, n in ((4, 60), (4, 50), (4, 40), (3, 60)):
    for env in ({}, {"TTN_MMA_DEEP": "16"}, {"TTN_MMA_DEEP": "0"}, {"TTN_MMA_RADIX": "0"}):
        for k in ("TTN_MMA_DEEP", "TTN_MMA_RADIX"):
            os.environ.pop(k, None)
        os.environ.update(env)
        s = t.continuous_siteinds(t.named_grid((n, 1)), base=base)
        f = t.rand_itn(s, link_space=16, rng=base, normalise=True)
        plan = f.plan()
        best = 1e9
        for _ in range(3):
            o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr())
            best = min(best, o.kernel_ms)
        print(f"base {base} n={n} {env}: {best:8.3f} ms {npts / best / 1e6:7.3f} G pts/s executed {o.flops_executed / npts:.0f}", flush=True)
