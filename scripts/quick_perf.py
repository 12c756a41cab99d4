"""Scratch perf probe (GPU): config-2 shape through the chain kernel, device-resident."""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import itna_b200 as t
from itna_b200 import _capi

npts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
dfma, dmma = C.c_double(), C.c_double()
_capi.check(_capi.lib().ttn_measure_fp64_peak(0, C.byref(dfma), C.byref(dmma)))
print(f"FP64 peak: DFMA {dfma.value:.2f} TF, DMMA {dmma.value:.2f} TF")

def run(name, f, ncol, kernel="auto", npts=npts):
    plan = f.plan()
    info = plan.info()
    x = torch.rand((npts, ncol), dtype=torch.float64, device="cuda:0")
    out = torch.empty(npts * (2 if info["is_complex"] else 1), dtype=torch.float64, device="cuda:0")
    best = 1e9
    for it in range(4):
        o = plan.evaluate_device(x.data_ptr(), npts, out.data_ptr(), kernel=kernel)
        best = min(best, o.kernel_ms)
    tf = info["flops_per_point"] * npts / (best * 1e-3) / 1e12
    print(f"{name:28s} kernel={_capi.KERNEL_NAMES[o.kernel_used]:8s} {npts:.1e} pts {best:9.3f} ms  "
          f"{npts / best / 1e3:9.2f} Mpts/s  {tf:6.2f} TFLOP/s ({100 * tf / dfma.value:5.1f}% of DFMA peak)")

g = t.named_comb_tree((2, 30))
s = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
f2 = t.rand_itn(s, link_space=16, rng=0, normalise=True)
run("cfg2 comb2x30 chi16", f2, 2, kernel="chain")
for spr in (1, 2, 3, 4):
    os.environ["TTN_MMA_SPR"] = str(spr)
    f2._plans.clear()
    run(f"cfg2 dmma spr={spr}", f2, 2, kernel="dmma")
del os.environ["TTN_MMA_SPR"]
s4 = t.continuous_siteinds(t.named_grid((28, 1)), map_dimension=2)
f4 = t.rand_itn(s4, link_space=32, rng=0, normalise=True)
for spr in (1, 2):
    os.environ["TTN_MMA_SPR"] = str(spr)
    f4._plans.clear()
    run(f"cfg4 dmma spr={spr}", f4, 2, kernel="dmma", npts=npts // 2)
del os.environ["TTN_MMA_SPR"]
s = t.continuous_siteinds(t.named_grid((28, 1)), map_dimension=2)
run("cfg4 mps28 chi32 chain", t.rand_itn(s, link_space=32, rng=0, normalise=True), 2, kernel="chain", npts=npts // 2)
s = t.continuous_siteinds(t.named_grid((20, 1)))
run("cfg1 sin qtt20 chi2 cplx", t.sin_itn(s, k=2.0), 1)
s = t.continuous_siteinds(t.named_grid((40, 1)), map_dimension=2)
f8 = t.rand_itn(s, link_space=8, rng=0, normalise=True)
run("mps40 chi8 chain", f8, 2, kernel="chain")
run("mps40 chi8 dmma", f8, 2, kernel="dmma")
run("exp product state chi1", t.exp_itn(s, k=1.0, dim=1), 2)
g = t.named_comb_tree((2, 30))
s = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
run("cfg2 generic kernel", t.rand_itn(s, link_space=16, rng=0, normalise=True), 2, kernel="generic", npts=min(npts, 1_000_000))
