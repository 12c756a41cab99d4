"""End-to-end rate from PINNED host buffers of chain shapes other than config 2 (A/B: TTN_HOST_QUANT=1 = doubles + light image)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import itna_b200 as t

N = 40_000_000
shapes = [("cfg4 shape: 28 sites chi=32 2-D", t.continuous_siteinds(t.named_grid((28, 1)), map_dimension=2), 32, 2),
          ("41 sites chi=16 1-D (41-bit runs: doubles)", t.continuous_siteinds(t.named_grid((41, 1))), 16, 1),
          ("comb 2x30 chi=8", t.continuous_siteinds(t.named_comb_tree((2, 30)), [[(i, j) for j in range(1, 31)] for i in (1, 2)]), 8, 2)]
for name, s, chi, nc in shapes:
    f = t.rand_itn(s, link_space=chi, rng=0, normalise=True)
    x = torch.rand((N, nc), dtype=torch.float64).pin_memory()
    out = torch.empty(N, dtype=torch.float64).pin_memory()
    xn, on = x.numpy(), out.numpy()
    for mode in ("1", None):
        if mode is None:
            os.environ.pop("TTN_HOST_QUANT", None)
        else:
            os.environ["TTN_HOST_QUANT"] = mode
        plan = f.plan()
        plan.evaluate_host(xn, out=on)
        t0 = time.perf_counter()
        for _ in range(3):
            _, o = plan.evaluate_host(xn, out=on)
        dt = (time.perf_counter() - t0) / 3
        print(f"{name:44s} TTN_HOST_QUANT={mode}: {N / dt / 1e9:6.3f} G points/s e2e, h2d {o.h2d_bytes / N:.1f} B/pt, staged {o.staged}", flush=True)
