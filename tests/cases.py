"""Network zoo shared by the parity tests: the reference's own test shapes
(test/test_realitensorfunction.jl, test/test_complexitensorfunction.jl) plus scaled-down
versions of the BASELINE.json configs."""
import numpy as np

import itna_b200 as t


def edge_points(L, ncol, rng, n_random=200):
    """Random points plus the edge cases of SURVEY §8(d): 0, 2^-L, 1-2^-L, 0.1, 0.675,
    nextafter(0.5, 0), 1.0 (saturates to all-ones), > 1."""
    d = 2.0 ** -L
    edge = [0.0, d, 1 - d, 0.1, 0.675, np.nextafter(0.5, 0), 1.0, 1.5, 0.25, 0.5, 0.625, 0.875,
            -0.0, 5.0 / 9.0, 1 - 2.0 ** -53]
    pts = rng.random((n_random, ncol))
    e = np.array([[edge[(i + 3 * j) % len(edge)] for j in range(ncol)] for i in range(len(edge))])
    return np.concatenate([pts, e], axis=0)


def real_cases():
    """(name, fitn, dims, L) — real-coordinate networks."""
    out = []
    # MPS, interleaved 2-D (examples/2d_laplace_solver.jl layout), chi=8
    g = t.named_grid((12, 1))
    s = t.continuous_siteinds(g, map_dimension=2)
    out.append(("mps2d_chi8", t.rand_itn(s, link_space=8, rng=1, normalise=True), [1, 2], 6))
    # comb tree 2 x 6: a chain once rooted at an end (BASELINE config 2 layout), chi=16
    g = t.named_comb_tree((2, 6))
    s = t.continuous_siteinds(g, [[(1, j) for j in range(1, 7)], [(2, j) for j in range(1, 7)]])
    out.append(("comb2x6_chi16", t.rand_itn(s, link_space=16, rng=2, normalise=True), [1, 2], 6))
    # comb tree 3 x 4 (a real tree: backbone of degree 3), chi=4
    g = t.named_comb_tree((3, 4))
    s = t.continuous_siteinds(g, [[(i, j) for j in range(1, 5)] for i in range(1, 4)])
    out.append(("comb3x4_chi4", t.rand_itn(s, link_space=4, rng=3, normalise=True), [1, 2, 3], 4))
    # binary tree depth 4, 3-D interleaved (BASELINE config 3 layout), chi=5; root has no site
    g = t.named_binary_tree(4)
    vs = g.vertices()
    with_site = vs[1:]
    dv = [with_site[i::3] for i in range(3)]
    s = t.continuous_siteinds(g, dv)
    out.append(("bintree4_chi5", t.rand_itn(s, link_space=5, rng=4, normalise=True), [1, 2, 3], 5))
    # binary tree depth 5, chi = 20 (BASELINE config 3 layout; selects the tree GEMM kernel, W = 32)
    g = t.named_binary_tree(5)
    with_site = g.vertices()[1:]
    s = t.continuous_siteinds(g, [with_site[i::3] for i in range(3)])
    out.append(("bintree5_chi20_tree", t.rand_itn(s, link_space=20, rng=14, normalise=True), [1, 2, 3], 10))
    # random labelled trees (test/test_realitensorfunction.jl:129), arbitrary degree
    for seed in (5, 6):
        g = t.uniform_tree(9, rng=seed).rename_vertices(lambda v: (v, 1))
        s = t.continuous_siteinds(g, map_dimension=2)
        out.append((f"unitree9_s{seed}", t.rand_itn(s, link_space=3, rng=seed, normalise=True), [1, 2], 5))
    # base 3 (test/test_realitensorfunction.jl:89-104)
    g = t.named_comb_tree((2, 3))
    s = t.continuous_siteinds(g, base=3)
    out.append(("base3_comb", t.rand_itn(s, link_space=3, rng=7, normalise=True), [1], 6))
    # base 3 MPS chain (chain kernel with 3 slices per vertex)
    g = t.named_grid((7, 1))
    s = t.continuous_siteinds(g, base=3)
    out.append(("base3_mps", t.rand_itn(s, link_space=4, rng=8, normalise=True), [1], 7))
    # complex-valued tensors with real coordinates (sin_itn is ComplexF64, config 1)
    g = t.named_grid((20, 1))
    s = t.continuous_siteinds(g)
    out.append(("sin_qtt20", t.sin_itn(s, k=3.0, a=0.25, c=0.8), [1], 20))
    # non-uniform link dimensions (after truncate / `+`)
    g = t.named_grid((10, 1))
    s = t.continuous_siteinds(g, map_dimension=2)
    f = t.rand_itn(s, link_space=3, rng=9, normalise=True) + t.cosh_itn(s, k=0.7, a=0.1, c=0.5, dim=2)
    out.append(("sum_chi3p2", f, [1, 2], 5))
    # chi = 32 MPS (BASELINE config 4 layout, shortened)
    g = t.named_grid((8, 1))
    s = t.continuous_siteinds(g, map_dimension=2)
    out.append(("mps2d_chi32", t.rand_itn(s, link_space=32, rng=10, normalise=True), [1, 2], 4))
    # wide chain: chi = 48 -> GEMM-regime kernel (width padded to 128)
    g = t.named_grid((6, 1))
    s = t.continuous_siteinds(g, map_dimension=2)
    out.append(("mps2d_chi48_gemm", t.rand_itn(s, link_space=48, rng=13, normalise=True), [1, 2], 3))
    # single vertex, two vertices
    g = t.named_grid((1, 1))
    s = t.continuous_siteinds(g)
    out.append(("single_vertex", t.rand_itn(s, link_space=1, rng=11), [1], 1))
    g = t.named_grid((2, 1))
    s = t.continuous_siteinds(g)
    out.append(("two_vertices", t.rand_itn(s, link_space=5, rng=12), [1], 2))
    return out


def complex_cases():
    """(name, fitn, dims, L) — ComplexIndexMap networks (coordinates are complex numbers)."""
    out = []
    # real and imaginary digits on different teeth (test/test_complexitensorfunction.jl:73-80)
    g = t.named_comb_tree((2, 3))
    s = t.complex_continuous_siteinds(g, [[(1, 1), (1, 2), (1, 3)]], [[(2, 1), (2, 2), (2, 3)]])
    out.append(("cplx_teeth", t.rand_itn(s, link_space=4, rng=20, eltype=complex, normalise=True), [1], 3))
    # alternating vertices (:121-129)
    g = t.named_grid((10, 1))
    s = t.complex_continuous_siteinds(g, [[(i, 1) for i in range(1, 11, 2)]], [[(i, 1) for i in range(2, 11, 2)]])
    out.append(("cplx_alt", t.rand_itn(s, link_space=6, rng=21, eltype=complex, normalise=True), [1], 5))
    # two site indices per vertex: Real of dim 1 + Imag of dim 2 (:183-207); BASELINE config 5 layout
    L = 10
    g = t.named_grid((L, 1))
    h = L // 2
    rv = [[(i, 1) for i in range(1, h + 1)], [(i, 1) for i in range(h + 1, L + 1)]]
    iv = [[(i, 1) for i in range(h + 1, L + 1)], [(i, 1) for i in range(1, h + 1)]]
    s = t.complex_continuous_siteinds(g, rv, iv)
    out.append(("cplx_2site", t.rand_itn(s, link_space=8, rng=22, eltype=complex, normalise=True), [1, 2], 5))
    # default complex map: every vertex carries Real+Imag of the same dimension (phys dim 4)
    g = t.named_grid((6, 1))
    s = t.complex_continuous_siteinds(g, map_dimension=2)
    out.append(("cplx_default2d", t.rand_itn(s, link_space=16, rng=23, eltype=complex, normalise=True), [1, 2], 3))
    # BASELINE config 5 layout (complex MPS, Real+Imag index per vertex, phys dim 4) at chi = 40:
    # real-embedded width 80 -> GEMM-regime kernel
    g = t.named_grid((5, 1))
    s = t.complex_continuous_siteinds(g, map_dimension=2)
    out.append(("cplx_cfg5_chi40_gemm", t.rand_itn(s, link_space=40, rng=25, eltype=complex, normalise=True), [1, 2], 3))
    # complex tree
    g = t.named_comb_tree((3, 3))
    s = t.complex_continuous_siteinds(g, map_dimension=3)
    out.append(("cplx_comb3x3", t.rand_itn(s, link_space=3, rng=24, eltype=complex, normalise=True), [1, 2, 3], 3))
    return out


def complex_points(L, ncol, rng, n_random=150):
    re = edge_points(L, ncol, rng, n_random)
    im = edge_points(L, ncol, np.random.default_rng(99), n_random)
    return re + 1j * im[::-1]


# ------------------------------------------------------------------ golden (reference known answers)
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_known_answers.json")


def load_golden():
    with open(GOLDEN) as fh:
        return json.load(fh)


def _cx(z):
    return complex(z[0], z[1])


def _num(z, want_complex):
    z = _cx(z)
    return z if (want_complex or z.imag != 0.0) else z.real


def build_siteinds(spec):
    kind, dims = spec["graph"]
    g = t.named_comb_tree(tuple(dims)) if kind == "comb" else t.named_grid(tuple(dims))
    base = spec.get("base", 2)
    tup = lambda vv: [[tuple(v) for v in verts] for verts in vv]
    if spec["map"] == "real":
        if "dimension_vertices" in spec:
            return t.continuous_siteinds(g, tup(spec["dimension_vertices"]), base=base)
        return t.continuous_siteinds(g, base=base, map_dimension=spec.get("map_dimension", 1))
    if "real_dimension_vertices" in spec:
        return t.complex_continuous_siteinds(g, tup(spec["real_dimension_vertices"]),
                                             tup(spec["imag_dimension_vertices"]), base=base)
    return t.complex_continuous_siteinds(g, base=base, map_dimension=spec.get("map_dimension", 1))


def build_golden_case(spec):
    """-> (fitn, point, dims, expected complex)"""
    s = build_siteinds(spec)
    cmap = spec["map"] == "complex"
    ctor = dict(exp=t.exp_itn, cosh=t.cosh_itn, sinh=t.sinh_itn, cos=t.cos_itn, sin=t.sin_itn,
                tanh=t.tanh_itn)
    f = None
    for term in spec["terms"]:
        if term["func"] == "const":
            ft = t.const_itn(s, c=_num(term["c"], False))
        else:
            kw = dict(k=_num(term["k"], False), a=_num(term["a"], False), c=_num(term["c"], False),
                      dim=term["dim"])
            if term["func"] == "tanh":
                kw["nterms"] = term["nterms"]
            ft = ctor[term["func"]](s, **kw)
        f = ft if f is None else f + ft
    point = [_cx(z) if cmap else z[0] for z in spec["point"]]
    return f, point, spec["dims"], _cx(spec["expected"])
