"""Pair-merging probe (GPU): merged vs one-vertex-per-position DMMA chain images; speed and agreement."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import itna_b200 as t
from itna_b200 import _capi

npts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
MERGES = tuple(os.environ.get("PROBE_MERGES", "1,2,3,4").split(","))

def run(name, f, ncol, npts=npts):
    res = {}
    for merge in MERGES:
        os.environ["TTN_MMA_MERGE"] = merge
        f._plans.clear()
        import time as _t
        _t0 = _t.time()
        plan = f.plan()
        _tb = _t.time() - _t0
        info = plan.info()
        torch.manual_seed(1)
        x = torch.rand((npts, ncol), dtype=torch.float64, device="cuda:0")
        out = torch.empty(npts * (2 if info["is_complex"] else 1), dtype=torch.float64, device="cuda:0")
        best = 1e9
        for it in range(4):
            o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr(), kernel="dmma")
            best = min(best, o.kernel_ms)
        res[merge] = out.clone()
        tf = info["flops_per_point"] * npts / (best * 1e-3) / 1e12
        print(f"{name:26s} merge={merge} {npts:.1e} pts {best:9.3f} ms {npts / best / 1e3:9.2f} Mpts/s  {tf:6.2f} TF algorithmic, "
              f"{o.flops_executed / (best * 1e-3) / 1e12:6.2f} TF executed, plan {_tb:.2f} s")
    a = res["1"]
    scale = a.abs().max().item()
    for m in MERGES[1:]:
        print(f"   merge={m}: max |merged - plain| / max|f| = {(a - res[m]).abs().max().item() / scale:.3e}")
    del os.environ["TTN_MMA_MERGE"]

g = t.named_comb_tree((2, 30))
s = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
run("cfg2 comb2x30 chi16", t.rand_itn(s, link_space=16, rng=0, normalise=True), 2)
s = t.continuous_siteinds(t.named_grid((40, 1)), map_dimension=2)
run("mps40 chi8", t.rand_itn(s, link_space=8, rng=0, normalise=True), 2)
s = t.continuous_siteinds(t.named_grid((41, 1)), map_dimension=1)
run("mps41 chi32 (v3 kernel)", t.rand_itn(s, link_space=32, rng=0, normalise=True), 1, npts // 2)
s = t.continuous_siteinds(t.named_grid((30, 1)), map_dimension=1)
run("mps30 chi16 complex", t.rand_itn(s, link_space=16, rng=0, eltype=complex, normalise=True), 1, npts // 2)
s = t.continuous_siteinds(t.named_grid((28, 1)), map_dimension=2)
run("cfg4 mps28 chi32 2-D", t.rand_itn(s, link_space=32, rng=0, normalise=True), 2, npts // 2)
