"""Small workload touching every kernel, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import itna_b200 as t
from itna_b200 import _capi
rng = np.random.default_rng(0)
nets = []
s = t.continuous_siteinds(t.named_grid((12, 1)), map_dimension=2)
nets.append(("mps chi16 (v6 / v5)", t.rand_itn(s, link_space=16, rng=1, normalise=True), 2, 3000))
nets.append(("mps chi32 (v6 / v3)", t.rand_itn(s, link_space=32, rng=2, normalise=True), 2, 1500))
nets.append(("mps chi48 (gemm)", t.rand_itn(s, link_space=48, rng=3, normalise=True), 2, 700))
s3 = t.continuous_siteinds(t.named_grid((9, 1)), base=3)
nets.append(("mps base3 chi8 (generic digits)", t.rand_itn(s3, link_space=8, rng=4, normalise=True), 1, 2000))
g = t.named_binary_tree(4)
ws = g.vertices()[1:]
sb = t.continuous_siteinds(g, [ws[i::3] for i in range(3)])
nets.append(("bintree chi20 (tree)", t.rand_itn(sb, link_space=20, rng=5, normalise=True), 3, 700))
s2 = t.continuous_siteinds(t.named_grid((30, 1)), map_dimension=2)
nets.append(("mps chi2 (table)", t.rand_itn(s2, link_space=2, rng=6, normalise=True), 2, 5000))
sc = t.complex_continuous_siteinds(t.named_grid((12, 1)), map_dimension=2)
nets.append(("complex mps chi2, 2 site indices per vertex (table)", t.rand_itn(sc, link_space=2, rng=7, eltype=complex, normalise=True), 4, 3000))
s90 = t.continuous_siteinds(t.named_grid((90, 1)), map_dimension=3)
nets.append(("mps90 chi2 3-D, two-word stream (table)", t.rand_itn(s90, link_space=2, rng=8, normalise=True), 3, 5000))
g40 = t.named_comb_tree((2, 40))
s40 = t.continuous_siteinds(g40, [[(i, j) for j in range(1, 41)] for i in (1, 2)])
nets.append(("comb2x40 chi1, 40-bit runs (table)", t.exp_itn(s40, k=-0.7, a=0.2, c=0.9, dim=2), 2, 5000))
# round 2: base-3 / base-4 chains on the team-sorted and table kernels (Digit4 K1), deep leaf / root tables built on
# the device (chains of > 32 bits), trees with merged single-child runs / multi-block classification / table-index
# gathers (real and complex combs), the refined (double-double) pass
s3l = t.continuous_siteinds(t.named_grid((24, 1)), base=3)
nets.append(("r2 mps24 base3 chi16 (team, Digit4)", t.rand_itn(s3l, link_space=16, rng=9, normalise=True), 1, 3000))
s4l = t.continuous_siteinds(t.named_grid((20, 1)), base=4, map_dimension=2)
nets.append(("r2 mps20 base4 chi2 (table, Digit4)", t.rand_itn(s4l, link_space=2, rng=10, normalise=True), 2, 3000))
s60 = t.continuous_siteinds(g40, [[(i, j) for j in range(1, 41)] for i in (1, 2)])
nets.append(("r2 comb2x40 chi16 (team, deep tables built on the device)", t.rand_itn(s60, link_space=16, rng=11, normalise=True), 2, 3000))
gc = t.named_comb_tree((3, 12))
sc3 = t.continuous_siteinds(gc, [[(j, i) for i in range(1, 13)] for j in range(1, 4)])
nets.append(("r2 comb3x12 chi16 (tree: tables, merged runs)", t.rand_itn(sc3, link_space=16, rng=12, normalise=True), 3, 3000))
scc = t.complex_continuous_siteinds(gc, [[(j, i) for i in range(1, 13)] for j in range(1, 4)],
                                    [[(j, i) for i in range(12, 0, -1)] for j in range(3, 0, -1)])
nets.append(("r2 complex comb3x12 chi6 (tree, complex fold)", t.rand_itn(scc, link_space=6, rng=13, eltype=complex, normalise=True), 6, 2000))
s3r = t.continuous_siteinds(t.named_grid((45, 1)), base=3)
nets.append(("r2 mps45 base3 chi16 (team, radix-3 group fields: 12 + 7 x 3 + 12 vertices)", t.rand_itn(s3r, link_space=16, rng=14, normalise=True), 1, 3000))
gs = t.NamedGraph([(i, 1) for i in range(6)], [((0, 1), (i, 1)) for i in range(1, 6)])
nets.append(("r2 star with 5 leaves chi2 (tree, binarised at plan time)", t.rand_itn(t.continuous_siteinds(gs, map_dimension=2), link_space=2, rng=15, normalise=True), 2, 3000))
skip = os.environ.get('SAN_SKIP', '')
only = os.environ.get('SAN_ONLY', '')
kernels_only = [k for k in os.environ.get('SAN_KERNELS', '').split(',') if k]   # e.g. SAN_KERNELS=dmma
# default plans: merged binary chains run the team-sorted kernel (v6); TTN_MMA_MERGE=1 keeps one vertex
# per position and exercises the warp-autonomous (v5) and warp-specialised (v3) kernels
for merge in (os.environ.get('SAN_MERGES', 'default,1').split(',')):
    if merge == 'default':
        os.environ.pop('TTN_MMA_MERGE', None)
    else:
        os.environ['TTN_MMA_MERGE'] = merge
    for name, f, ncol, n in nets:
        if (skip and skip in name) or (only and only not in name):
            continue
        f._plans.clear()
        plan = f.plan()
        pts = rng.random((n, ncol))
        avail = plan.info()["kernels_available"]
        for kname, kid in _capi.KERNEL_IDS.items():
            if kernels_only and kname not in kernels_only:
                continue
            if kid and kname != 'grid' and avail & (1 << kid) and (merge == 'default' or kname == 'dmma'):
                out, o = plan.evaluate_host(pts, kernel=kname, reduce_sum=True)
                print(f"merge={merge}", name, kname, "ok", float(np.abs(out).max()))
        plan.digits_host(pts)
        if name.startswith("r2") and merge == 'default':
            out, o = plan.evaluate_host(pts, accuracy="refined", reduce_sum=True)
            print(name, "refined ok", int(o.n_refined))
        f._plans.clear()
if not only or only == 'r2':
    # host-side quantisation of pageable run-path coordinates (pack_coords -> CoordSource::qcoords), >= 2^20 points
    s2 = t.continuous_siteinds(g40, [[(i, j) for j in range(1, 41)] for i in (1, 2)])
    s2b = t.continuous_siteinds(t.named_comb_tree((2, 30)), [[(i, j) for j in range(1, 31)] for i in (1, 2)])
    fq = t.rand_itn(s2b, link_space=16, rng=16, normalise=True)
    out, o = fq.plan().evaluate_host(rng.random((1_100_000, 2)), reduce_sum=True)
    print("r2 quantised pageable coordinates ok", int(o.staged), int(o.h2d_bytes))
print("done")
