"""CPU audit of the table kernel's plan-time images at scale: walk the group tables (numpy, vectorised — the same
products the kernel forms, one fused multiply-add apart) for N random points of the small-chi bench networks and
compare with the 80-bit contraction of the packed tensors.  No GPU needed (debug hook ttn_debug_table_image).
    python scripts/table_accuracy.py [n_points=200000]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import itna_b200 as t
import oracle as orc
from test_table_image import table_image

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000


def walk_vec(im, dig):
    H, E = im["H"], 2 if im["cplx"] else 1
    G = len(im["gbits"])
    w = np.zeros(len(dig), dtype=object)
    for site in range(dig.shape[1]):
        w = w + (dig[:, site].astype(object) << int(im["bitpos"][site]))

    def entries(g, s, shape):
        nd = int(np.prod(shape)) * E
        img, off = im["image"], im["goff"][g]
        if im["rep"]:
            if nd == 1:
                raw = img[off + s * 16][:, None]
            else:
                c = nd // 2
                idx = off + (s[:, None, None] * c + np.arange(c)[None, :, None]) * 16 + np.arange(2)[None, None, :]
                raw = img[idx].reshape(len(s), nd)
        else:
            c = nd // 2
            if c >= 2:
                sw = {2: (s >> 2) & 1, 4: (s >> 1) & 3, 8: s & 7}[c]
                pos = (np.arange(c)[None, :] ^ sw[:, None])            # chunk j stored at position j ^ sw
                idx = off + s[:, None, None] * nd + pos[:, :, None] * 2 + np.arange(2)[None, None, :]
                raw = img[idx].reshape(len(s), nd)
            else:
                raw = img[off + s[:, None] * nd + np.arange(nd)[None, :]]
        if E == 2:
            raw = raw[:, 0::2] + 1j * raw[:, 1::2]
        return raw.reshape((len(s),) + tuple(shape))

    def take(bits):
        nonlocal w
        s = np.array([int(x & ((1 << bits) - 1)) for x in w], dtype=np.int64)
        w = np.array([x >> bits for x in w], dtype=object)
        return s

    v = entries(0, take(im["gbits"][0]), (H,))
    for g in range(1, G - 1):
        v = np.einsum("pi,pij->pj", v, entries(g, take(im["gbits"][g]), (H, H)))
    return np.einsum("pi,pi->p", v, entries(G - 1, take(im["gbits"][G - 1]), (H,)))


g = t.named_comb_tree((2, 30))
s2 = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
nets = [("exp chi1 2x30 (bench --config 6)", t.exp_itn(s2, k=0.9, a=0.1, c=1.2, dim=1)),
        ("rand chi2 2x30 (bench --config 7)", t.rand_itn(s2, link_space=2, rng=20267, normalise=True)),
        ("rand chi4 2x30", t.rand_itn(s2, link_space=4, rng=20268, normalise=True)),
        ("sin QTT 20 bits (config 1)", t.sin_itn(t.continuous_siteinds(t.named_grid((20, 1))), k=2.0, a=0.3, c=1.1))]
rng = np.random.default_rng(7)
for name, f in nets:
    packed = t.pack(f)
    im = table_image(packed)
    pts = rng.random((n, packed.n_coords))
    dig = orc.digits(packed, pts)
    ref = orc.evaluate(packed, pts, orc.ORACLE_LD, nthreads=orc.max_threads())
    f64 = orc.evaluate(packed, pts, orc.ORACLE_F64, nthreads=orc.max_threads())
    got = walk_vec(im, dig)
    e_t, e_f = orc.error_metric(got, ref), orc.error_metric(f64, ref)
    print(f"{name:34s} {n} points, groups {im['gbits']} rep={im['rep']}: tables max {e_t.max():.2e} p99.9 {np.quantile(e_t, 0.999):.2e} | "
          f"per-vertex FP64 max {e_f.max():.2e} p99.9 {np.quantile(e_f, 0.999):.2e}")
