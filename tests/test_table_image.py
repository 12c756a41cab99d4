"""CPU test of the table kernel's plan-time image (k_chain_table.cu: make_table_image) — the host logic
that decides what the HBM-bound kernel computes.  The image (group tables of pre-contracted vertices) is
fetched through the debug hook ttn_debug_table_image (no CUDA calls) and walked in numpy exactly as the
kernel walks it: packed slice stream -> leaf vector, middle matrices, root vector.  Checked against the
80-bit oracle contraction of the same packed tensors."""
import ctypes as C

import numpy as np
import pytest

import cases
import itna_b200 as t
import oracle as orc
from itna_b200 import _capi

MAXG = 16


def table_image(packed, budget_kb=200):
    L = _capi.lib()
    meta = (C.c_int32 * (8 + 2 * MAXG))()
    cap = 1 << 18
    img = np.zeros(cap, dtype=np.float64)
    bitpos = np.zeros(max(len(packed.site_dim), 1), dtype=np.int32)
    rc = L.ttn_debug_table_image(C.byref(packed.desc()), budget_kb, meta, img.ctypes.data_as(C.c_void_p),
                                 C.c_int64(cap), bitpos.ctypes.data_as(C.c_void_p))
    assert rc == 0
    meta = list(meta)
    if not meta[0]:
        return None
    G = meta[4]
    return dict(H=meta[1], cplx=meta[2], bits0=meta[3], rep=meta[6], gbits=meta[8:8 + G], goff=meta[8 + MAXG:8 + MAXG + G],
                image=img[:meta[5]].copy(), bitpos=bitpos)


def walk(im, digits):
    """What chain_table_kernel does per point, in numpy (python ints for the 128-bit stream)."""
    H, E = im["H"], 2 if im["cplx"] else 1
    G = len(im["gbits"])

    def entry(g, s, shape):
        n = int(np.prod(shape)) * E
        if im["rep"]:                    # replicated layout: chunk j of entry s = line s C + j, identical copies
            if n == 1:
                line = im["image"][im["goff"][g] + s * 16: im["goff"][g] + (s + 1) * 16]
                assert (line == line[0]).all()
                return line[:1].reshape(shape)
            c = n // 2
            lines = im["image"][im["goff"][g] + s * c * 16: im["goff"][g] + (s + 1) * c * 16].reshape(c, 8, 2)
            assert (lines == lines[:, :1, :]).all()
            raw = lines[:, 0, :].reshape(-1)
            return (raw[0::2] + 1j * raw[1::2]).reshape(shape) if E == 2 else raw.reshape(shape)
        raw = im["image"][im["goff"][g] + s * n: im["goff"][g] + (s + 1) * n]
        c = n // 2                       # 16-byte chunks; chunk j is stored at position j ^ swizzle(s)
        if c >= 2:
            sw = {2: (s >> 2) & 1, 4: (s >> 1) & 3, 8: s & 7}[c]
            raw = raw.reshape(c, 2)[[j ^ sw for j in range(c)]].reshape(-1)
        return (raw[0::2] + 1j * raw[1::2]).reshape(shape) if E == 2 else raw.reshape(shape)

    out = []
    for row in digits:
        w = 0
        for site, v in enumerate(row):
            w += int(v) << int(im["bitpos"][site])
        s = w & ((1 << im["gbits"][0]) - 1)
        w >>= im["gbits"][0]
        vec = entry(0, s, (H,))
        for g in range(1, G - 1):
            s = w & ((1 << im["gbits"][g]) - 1)
            w >>= im["gbits"][g]
            vec = vec @ entry(g, s, (H, H))
        s = w & ((1 << im["gbits"][G - 1]) - 1)
        w >>= im["gbits"][G - 1]
        assert w == 0
        out.append(vec @ entry(G - 1, s, (H,)))
    return np.array(out)


def narrow_cases():
    out = []
    s1 = t.continuous_siteinds(t.named_grid((20, 1)))
    out.append(("sin_qtt20", t.sin_itn(s1, k=3.0, a=0.25, c=0.8), 200))              # complex chi 2, 20 bits
    out.append(("exp_1d", t.exp_itn(s1, k=-1.3, a=0.2, c=1.7), 200))                 # chi 1
    g = t.named_comb_tree((2, 30))
    s2 = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
    out.append(("exp_comb2x30", t.exp_itn(s2, k=0.9, a=0.1, c=1.2, dim=1) , 200))    # bench shape, chi 1
    out.append(("cosh_comb2x30", t.cosh_itn(s2, k=0.9, a=0.1, c=1.2, dim=2), 200))   # chi 2 real, 60 bits
    out.append(("rand_chi2_comb2x30", t.rand_itn(s2, link_space=2, rng=5, normalise=True), 200))
    out.append(("rand_chi2_comb2x30_100k", t.rand_itn(s2, link_space=2, rng=5, normalise=True), 100))
    s3 = t.continuous_siteinds(t.named_grid((24, 1)), map_dimension=3)
    out.append(("rand_chi4_mps3d", t.rand_itn(s3, link_space=4, rng=6, normalise=True), 200))
    out.append(("rand_chi3_mps3d", t.rand_itn(s3, link_space=3, rng=7, normalise=True), 64))
    sc = t.complex_continuous_siteinds(t.named_grid((12, 1)), map_dimension=2)       # Real+Imag index per vertex
    out.append(("cplx_2site_chi2", t.rand_itn(sc, link_space=2, rng=8, eltype=complex, normalise=True), 200))
    out.append(("cplx_2site_chi1", t.rand_itn(sc, link_space=1, rng=9, eltype=complex, normalise=True), 16))
    sa = t.complex_continuous_siteinds(t.named_grid((10, 1)), [[(i, 1) for i in range(1, 11, 2)]],
                                       [[(i, 1) for i in range(2, 11, 2)]])
    out.append(("cplx_alt_chi2", t.rand_itn(sa, link_space=2, rng=10, eltype=complex, normalise=True), 8))
    # vertices without a site index inside the chain
    sgap = t.continuous_siteinds(t.named_grid((9, 1)), [[(1, 1), (4, 1), (7, 1)], [(2, 1), (8, 1)]])
    out.append(("rand_chi3_siteless", t.rand_itn(sgap, link_space=3, rng=3, normalise=True), 200))
    # more than 64 slice bits
    g40 = t.named_comb_tree((2, 40))
    s40 = t.continuous_siteinds(g40, [[(i, j) for j in range(1, 41)] for i in (1, 2)])
    out.append(("rand_chi2_comb2x40", t.rand_itn(s40, link_space=2, rng=11, normalise=True), 200))
    s90 = t.continuous_siteinds(t.named_grid((90, 1)), map_dimension=3)
    out.append(("exp_mps90_3d", t.exp_itn(s90, k=0.4, a=-0.1, c=1.3, dim=3), 200))
    sc40 = t.complex_continuous_siteinds(t.named_grid((40, 1)), map_dimension=1)
    out.append(("cplx_2site_chi2_40", t.rand_itn(sc40, link_space=2, rng=14, eltype=complex, normalise=True), 200))
    # chi = 1 with the plain (not replicated) layout: negative budget
    out.append(("exp_comb2x30_plain", t.exp_itn(s2, k=0.9, a=0.1, c=1.2, dim=1), -200))
    out.append(("cplx_2site_chi1_plain", t.rand_itn(sc, link_space=1, rng=9, eltype=complex, normalise=True), -16))
    return out


@pytest.mark.parametrize("case", narrow_cases(), ids=lambda c: c[0])
def test_table_image_walk_matches_oracle(case):
    name, f, kb = case
    packed = t.pack(f)
    im = table_image(packed, kb)
    assert im is not None, name
    assert sum(im["gbits"]) == packed.n_vertices * im["bits0"]
    assert im["image"].size * 8 <= abs(kb) * 1024 and all(o % 16 == 0 for o in im["goff"])
    assert im["rep"] in (0, 1) and (kb > 0 or im["rep"] == 0), (name, im["gbits"])
    rng = np.random.default_rng(3)
    nc = packed.n_coords
    pts = np.concatenate([rng.random((300 if packed.n_vertices <= 60 else 60, nc)), cases.edge_points(20, nc, rng, 0)])
    dig = orc.digits(packed, pts)
    ref = orc.evaluate(packed, pts, orc.ORACLE_LD)
    got = walk(im, dig)
    assert orc.error_metric(got, ref).max() < 1e-12, name


def test_table_image_not_for_wide_or_nonbinary_or_trees():
    s = t.continuous_siteinds(t.named_grid((12, 1)), map_dimension=2)
    assert table_image(t.pack(t.rand_itn(s, link_space=8, rng=1))) is None            # chi 8
    s3 = t.continuous_siteinds(t.named_grid((7, 1)), base=3)
    assert table_image(t.pack(t.rand_itn(s3, link_space=2, rng=1))) is None           # base 3
    g = t.named_comb_tree((3, 4))
    sc = t.continuous_siteinds(g, [[(i, j) for j in range(1, 5)] for i in range(1, 4)])
    assert table_image(t.pack(t.rand_itn(sc, link_space=2, rng=1))) is None           # a real tree
    sz = t.complex_continuous_siteinds(t.named_grid((6, 1)), map_dimension=2)
    assert table_image(t.pack(t.rand_itn(sz, link_space=3, rng=1, eltype=complex))) is None  # complex chi 3
