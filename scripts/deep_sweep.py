"""Time per point vs number of middle rounds with deep leaf/root tables (chi = 16 binary chains)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itna_b200 as t
npts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 40_000_000
LS = tuple(int(a) for a in os.environ.get('SWEEP_L', '40,44,48,52,56,60').split(','))
for L in LS:
    s = t.continuous_siteinds(t.named_grid((L, 1)), map_dimension=1)
    f = t.rand_itn(s, link_space=16, rng=0, normalise=True)
    plan = f.plan()
    x = torch.rand((npts, 1), dtype=torch.float64, device="cuda:0")
    out = torch.empty(npts, dtype=torch.float64, device="cuda:0")
    best = 1e9
    for _ in range(4):
        o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr(), kernel="dmma")
        best = min(best, o.kernel_ms)
    print(f"L={L}: rounds {plan.info()["n_vertices"] and int(round((o.flops_executed / npts - 32) / 512))}  {best:8.3f} ms  {npts / best / 1e6:8.2f} G pts/s  executed {o.flops_executed / npts:.0f} flop/pt")
