"""Accuracy of the merged DMMA chain images against the 80-bit oracle (GPU + oracle)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import itna_b200 as t
import oracle as orc

g = t.named_comb_tree((2, 30))
dv = [[(i, j) for j in range(1, 31)] for i in (1, 2)]
s = t.continuous_siteinds(g, dv)
f1 = t.rand_itn(s, link_space=16, rng=7, normalise=True)
f2 = t.rand_itn(s, link_space=16, rng=8, normalise=True)
rng = np.random.default_rng(3)
pts = rng.random((200_000, 2))
packed = f1.plan().packed
ref = orc.evaluate(packed, pts, orc.ORACLE_LD, nthreads=orc.max_threads())
bp = orc.evaluate(packed, pts, orc.ORACLE_BP, nthreads=orc.max_threads())
f64 = orc.evaluate(packed, pts, orc.ORACLE_F64, nthreads=orc.max_threads())
for name, v in (("oracle BP", bp), ("oracle F64", f64)):
    e = orc.error_metric(v, ref)
    print(f"{name:12s} p50 {np.quantile(e, .5):.2e} p99.9 {np.quantile(e, .999):.2e} max {e.max():.2e}")
f12 = f1 + f2
ref12 = orc.evaluate(f12.plan().packed, pts, orc.ORACLE_LD, nthreads=orc.max_threads())
for m in ("1", "2", "3", "4"):
    os.environ["TTN_MMA_MERGE"] = m
    for f in (f1, f2, f12): f._plans.clear()
    a, b, c = t.evaluate(f1, pts, kernel="dmma"), t.evaluate(f2, pts, kernel="dmma"), t.evaluate(f12, pts, kernel="dmma")
    e = orc.error_metric(a, ref)
    e12 = orc.error_metric(c, ref12)
    lin = orc.error_metric(c, a + b)
    print(f"merge={m}: f1 p50 {np.quantile(e, .5):.2e} p99.9 {np.quantile(e, .999):.2e} max {e.max():.2e} | "
          f"f1+f2 (chi 32) p99.9 {np.quantile(e12, .999):.2e} max {e12.max():.2e} | linearity max {lin.max():.2e}")
    i = int(np.argmax(lin))
    print(f"     worst linearity point: a={a[i]:.6e} b={b[i]:.6e} c={c[i]:.6e} ref12={ref12[i]:.6e} a+b={a[i]+b[i]:.6e}")
