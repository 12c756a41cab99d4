"""Tree-topology probe (GPU): 3-tooth comb trees (examples/construct_multi_dimensional_function.jl layout), real and
complex, small and moderate chi, through the planner's choice.  Prints kernel, ms, points/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itna_b200 as t
from itna_b200 import _capi

npts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
L = 20
g = t.named_comb_tree((3, L))
rv = [[(j, i) for i in range(1, L + 1)] for j in range(1, 4)]
iv = [[(j, i) for i in range(L, 0, -1)] for j in range(3, 0, -1)]
cases = []
for chi in (2, 4, 8, 16):
    s = t.continuous_siteinds(g, rv)
    cases.append((f"real comb 3x{L} chi={chi}", t.rand_itn(s, link_space=chi, rng=chi, normalise=True), 3))
for chi in (2, 4, 8):
    s = t.complex_continuous_siteinds(g, rv, iv)
    cases.append((f"complex comb 3x{L} chi={chi}", t.rand_itn(s, link_space=chi, rng=chi, eltype=complex, normalise=True), 6))
for name, f, ncol in cases:
    plan = f.plan()
    info = plan.info()
    n = npts
    x = torch.rand((n, ncol), dtype=torch.float64, device="cuda:0")
    out = torch.empty(n * (2 if info["is_complex"] else 1), dtype=torch.float64, device="cuda:0")
    best = 1e9
    for it in range(3):
        o = plan.evaluate_device(x.data_ptr(), n, out.data_ptr())
        best = min(best, o.kernel_ms)
    print(f"{name:28s} kernel={_capi.KERNEL_NAMES[o.kernel_used]:8s} {n:.1e} pts {best:9.3f} ms {n / best / 1e3:9.1f} M pts/s "
          f"{info['flops_per_point'] * n / best / 1e9:7.2f} TF (rule)", flush=True)
