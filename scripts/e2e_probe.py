"""End-to-end (pinned host buffers) chunk-size probe on the cfg2 shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import itna_b200 as t
npts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
g = t.named_comb_tree((2, 30))
s = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
f = t.rand_itn(s, link_space=16, rng=0, normalise=True)
plan = f.plan()
x = torch.rand((npts, 2), dtype=torch.float64).pin_memory()
o = torch.empty(npts, dtype=torch.float64).pin_memory()
xn, on = x.numpy(), o.numpy()
for chunk in (1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
    for _ in range(2):
        plan.evaluate_host(xn, out=on, chunk_points=chunk)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        plan.evaluate_host(xn, out=on, chunk_points=chunk)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"chunk 2^{chunk.bit_length() - 1}: {dt * 1e3:8.2f} ms/step  {npts / dt / 1e9:6.3f} G pts/s  H2D {npts * 16 / dt / 1e9:5.1f} GB/s")
