"""generic vs tree kernel (with subtree tables) on small-chi trees."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import itna_b200 as t
from itna_b200 import _capi
npts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
cases = []
for chi in (2, 3, 4, 6, 8, 12):
    g = t.named_comb_tree((3, 20))
    s = t.continuous_siteinds(g, [[(i, j) for j in range(1, 21)] for i in (1, 2, 3)])
    cases.append((f"comb3x20 chi{chi}", t.rand_itn(s, link_space=chi, rng=1, normalise=True), 3))
    g = t.named_binary_tree(5)
    ws = g.vertices()
    s = t.continuous_siteinds(g, [ws[i::2] for i in range(2)])
    cases.append((f"bintree5 chi{chi}", t.rand_itn(s, link_space=chi, rng=2, normalise=True), 2))
for name, f, ncol in cases:
    plan = f.plan()
    info = plan.info()
    x = torch.rand((npts, ncol), dtype=torch.float64, device="cuda:0")
    out = torch.empty(npts, dtype=torch.float64, device="cuda:0")
    line = f"{name:18s} auto={_capi.KERNEL_NAMES[info['auto_kernel']]:8s}"
    for k in ("generic", "tree"):
        if not info["kernels_available"] & (1 << _capi.KERNEL_IDS[k]):
            continue
        n = npts if k == "tree" else min(npts, 1_000_000)
        best = 1e9
        for _ in range(3):
            o = plan.evaluate_device(x.data_ptr(), n, out.data_ptr(), kernel=k)
            best = min(best, o.kernel_ms)
        line += f"  {k}: {n / best / 1e3:9.2f} Mpts/s"
    print(line)
