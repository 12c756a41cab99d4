import os, sys
sys.path.insert(0, os.getcwd())
import torch
import itna_b200 as t
npts = 40_000_000
x = torch.rand((npts, 2), dtype=torch.float64, device="cuda:0")
out = torch.empty(npts, dtype=torch.float64, device="cuda:0")
def run(name, f, ncol, env):
    for k in ("TTN_MMA_DEEP", "TTN_MMA_RADIX", "TTN_MMA_MERGE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    f.invalidate_plans()
    plan = f.plan()
    best = 1e9
    xx = x if ncol == 2 else x[:, :1].contiguous()
    for _ in range(3):
        o = plan.evaluate_device(xx.data_ptr(), npts, out.data_ptr())
        best = min(best, o.kernel_ms)
    print(f"{name} {env}: {best:8.3f} ms executed {o.flops_executed / npts:.0f} flop/pt", flush=True)
for n in (60, 120):
    g = t.named_comb_tree((2, n // 2))
    s = t.continuous_siteinds(g, [[(i, j) for j in range(1, n // 2 + 1)] for i in (1, 2)])
    f = t.rand_itn(s, link_space=16, rng=1, normalise=True)
    run(f"base2 2x{n//2} run-path", f, 2, {"TTN_MMA_DEEP": "0"})
    run(f"base2 2x{n//2} run-path", f, 2, {"TTN_MMA_DEEP": "0", "TTN_MMA_MERGE": "2"})
# interleaved binary chain: no run path in the team kernel (Digit2 table loop)
s = t.continuous_siteinds(t.named_grid((60, 1)), map_dimension=2)
f = t.rand_itn(s, link_space=16, rng=2, normalise=True)
run("base2 interleaved 60 (Digit2 loop)", f, 2, {"TTN_MMA_DEEP": "0"})
s = t.continuous_siteinds(t.named_grid((60, 1)), base=4)
f = t.rand_itn(s, link_space=16, rng=4, normalise=True)
run("base4 60", f, 1, {"TTN_MMA_DEEP": "0"})
s = t.continuous_siteinds(t.named_grid((30, 1)), base=4)
f = t.rand_itn(s, link_space=16, rng=4, normalise=True)
run("base4 30", f, 1, {"TTN_MMA_DEEP": "0"})
