"""Write the bench network (BASELINE configs[1]: comb tree 2x30, chi = 16, the very network bench.py times) or
another config's network with N points to a portable file for julia/ref_evaluate.jl:
    python scripts/export_bench_network.py [config=2] [n_points=1000] [out=bench_cfg2.ttn.json]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import itna_b200 as t
import bench

config = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
out = sys.argv[3] if len(sys.argv) > 3 else f"bench_cfg{config}.ttn.json"
f, ncol, _, desc = bench.build_workload(config)
rng = np.random.default_rng(2026)
pts = rng.random((n, ncol))
if isinstance(f.indexmap, t.ComplexIndexMap):
    pts = pts[:, 0::2] + 1j * pts[:, 1::2]
t.save_ttn(f, out, points=pts)
print(f"wrote {out}: {desc}; {n} points; {os.path.getsize(out) / 1e6:.2f} MB")
