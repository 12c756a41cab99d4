"""Subtree tables on/off on the config-3 tree: bitwise agreement and error vs the 80-bit oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import itna_b200 as t
import oracle as orc
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = t.named_binary_tree(7)
vs = g.vertices()
s = t.continuous_siteinds(g, [vs[7:][i::3] for i in range(3)])
f = t.rand_itn(s, link_space=chi, rng=3, normalise=True)
rng = np.random.default_rng(5)
pts = rng.random((4096, 3))
res = {}
for bits in ("0", "16"):
    os.environ["TTN_TREE_TABLE_BITS"] = bits
    f._plans.clear()
    res[bits] = t.evaluate(f, pts, kernel="tree")
print("bitwise equal:", bool((res["0"] == res["16"]).all()), " max |diff|/max|f|:", np.abs(res["0"] - res["16"]).max() / np.abs(res["0"]).max())
ref = orc.evaluate(f.plan().packed, pts[:512], orc.ORACLE_LD, nthreads=orc.max_threads())
f64 = orc.evaluate(f.plan().packed, pts[:512], orc.ORACLE_F64, nthreads=orc.max_threads())
for k in res:
    e = orc.error_metric(res[k][:512], ref)
    print(f"tables bits={k}: p50 {np.quantile(e, .5):.2e} p99 {np.quantile(e, .99):.2e} max {e.max():.2e}")
e = orc.error_metric(f64, ref)
print(f"oracle F64     : p50 {np.quantile(e, .5):.2e} p99 {np.quantile(e, .99):.2e} max {e.max():.2e}")
