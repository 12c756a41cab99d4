"""CPU tests (no GPU) of two host-side transformations of libttneval.so, through test hooks that make no CUDA call:
 * pack_coords — the quantisation of run-path coordinates the staging threads apply to pageable arrays
   (TTN_STAGE_AUTO): must give exactly the digits of the reference's greedy loop (abstractindexmap.jl:121-138);
 * binarize_desc — the plan-time splitting of vertices with more than two children into Kronecker pair vertices: the
   binarised description must define the same function (checked with the CPU oracle on both descriptions)."""
import ctypes as C

import numpy as np

import itna_b200 as t
import oracle as orc
from itna_b200 import _capi
from itna_b200.packer import PackedNetwork


def test_host_quantisation_matches_the_greedy_digits():
    L = _capi.lib()
    fn = L.ttn_debug_pack_coords
    fn.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    fn.restype = C.c_int
    rng = np.random.default_rng(0)
    for nc, Ls in ((2, (30, 30)), (1, (20,)), (2, (31, 7)), (3, (12, 32, 25)), (4, (10, 11, 12, 13))):
        n = 200_003
        x = rng.random((n, nc))
        js = np.arange(1, 33)
        edge = np.concatenate([2.0 ** -js, np.nextafter(2.0 ** -js, 0), np.nextafter(2.0 ** -js, 1), 1 - 2.0 ** -js,
                               [0.0, 1.0, 1.5, 1e300, 5e-324, 1 - 2.0 ** -53, np.nextafter(1.0, 0)]])
        x[: len(edge), :] = np.stack([np.roll(edge, 3 * c) for c in range(nc)], axis=1)
        q = np.zeros((n, nc), dtype=np.uint32)
        Larr = np.asarray(Ls, dtype=np.int32)
        rc = fn(x.ctypes.data_as(C.c_void_p), n, nc, Larr.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p))
        assert rc == 0
        for c in range(nc):
            # the greedy loop on base-2 thresholds 2^-k, k = 1..L (exact arithmetic in FP64): digit k = bit L-k of q
            xr = x[:, c].copy()
            want = np.zeros(n, dtype=np.uint64)
            for k in range(1, Ls[c] + 1):
                ge = xr >= 2.0 ** -k
                xr = np.where(ge, xr - 2.0 ** -k, xr)
                want |= ge.astype(np.uint64) << np.uint64(Ls[c] - k)
            assert (q[:, c].astype(np.uint64) == want).all(), (nc, Ls, c)
    # domain errors are flagged by the host pass
    bad = rng.random((100_000, 2))
    Larr = np.asarray((30, 30), dtype=np.int32)
    q = np.zeros((100_000, 2), dtype=np.uint32)
    for v in (-1e-300, np.nan, -np.inf):
        bad[77_777, 1] = v
        assert fn(bad.ctypes.data_as(C.c_void_p), 100_000, 2, Larr.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p)) == 1


def _binarized(packed):
    L = _capi.lib()
    fn = L.ttn_debug_binarize
    fn.argtypes = [C.POINTER(_capi.ttn_desc), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    fn.restype = C.c_int
    cap = 4 * packed.n_vertices
    parent, link = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
    tptr = np.zeros(cap + 1, dtype=np.int64)
    nc = 2 if packed.is_complex else 1
    tcap = int(packed.tensors.size * (1 if packed.tensors.dtype != np.complex128 else 2)) + cap * 64 * 64 * nc
    tens = np.zeros(tcap, dtype=np.float64)
    n = fn(C.byref(packed.desc()), cap, parent.ctypes.data_as(C.c_void_p), link.ctypes.data_as(C.c_void_p),
           tptr.ctypes.data_as(C.c_void_p), tens.ctypes.data_as(C.c_void_p), tcap)
    assert n >= 0
    if n == 0:
        return None
    site_ptr = np.concatenate([packed.site_ptr, np.full(n - packed.n_vertices, packed.site_ptr[-1], dtype=np.int32)])
    tensors = tens[: tptr[n] * nc].copy()
    if packed.is_complex:
        tensors = tensors.view(np.complex128)
    kw = dict(packed.__dict__)
    kw.pop("_desc", None)
    kw.update(n_vertices=n, parent=parent[:n].copy(), link_dim=link[:n].copy(), site_ptr=site_ptr.astype(np.int32),
              tensor_ptr=tptr[: n + 1].copy(), tensors=tensors)
    return PackedNetwork(**kw)


def test_binarised_description_defines_the_same_function():
    rng = np.random.default_rng(1)
    nets = []
    for k, chi in ((4, 3), (5, 2), (3, 8)):
        g = t.NamedGraph([(i, 1) for i in range(k + 1)], [((0, 1), (i, 1)) for i in range(1, k + 1)])
        nets.append((f"star{k}", t.rand_itn(t.continuous_siteinds(g, map_dimension=2), link_space=chi, rng=k, normalise=True)))
    verts = [(i, 1) for i in range(5)] + [(i, 2) for i in range(1, 4)] + [(i, 3) for i in range(1, 4)]
    edges = [((i, 1), (i + 1, 1)) for i in range(4)] + [((i, 1), (i, 2)) for i in range(1, 4)] + [((i, 1), (i, 3)) for i in range(1, 4)]
    g = t.NamedGraph(verts, edges)
    nets.append(("caterpillar", t.rand_itn(t.continuous_siteinds(g, map_dimension=3), link_space=4, rng=9, normalise=True)))
    nets.append(("caterpillar_cplx", t.rand_itn(t.complex_continuous_siteinds(g, map_dimension=2), link_space=3, rng=10,
                                                eltype=complex, normalise=True)))
    for seed in (6, 11, 12):
        g = t.uniform_tree(14, rng=seed).rename_vertices(lambda v: (v, 1))
        nets.append((f"unitree14_s{seed}", t.rand_itn(t.continuous_siteinds(g, map_dimension=2), link_space=3, rng=seed, normalise=True)))
    seen_split = 0
    for name, f in nets:
        packed = t.pack(f)
        b = _binarized(packed)
        children = np.bincount(packed.parent[packed.parent >= 0], minlength=packed.n_vertices)
        if children.max() <= 2:
            assert b is None, name                    # nothing to split
            continue
        assert b is not None, name
        seen_split += 1
        assert b.n_vertices == packed.n_vertices + int(np.maximum(children - 2, 0).sum())
        assert np.bincount(b.parent[b.parent >= 0], minlength=b.n_vertices).max() <= 2
        assert (b.tensor_ptr[: packed.n_vertices + 1] == packed.tensor_ptr).all()       # original tensors untouched
        coords = rng.random((300, packed.n_coords))
        ref = orc.evaluate(packed, coords, orc.ORACLE_LD)
        got = orc.evaluate(b, coords, orc.ORACLE_LD)
        assert orc.error_metric(got, ref).max() < 1e-15, name
        assert (orc.digits(b, coords) == orc.digits(packed, coords)).all()
    assert seen_split >= 5
    # a chain and a binary tree are left alone; a star whose Kronecker products would exceed width 64 too
    for f in (t.rand_itn(t.continuous_siteinds(t.named_grid((8, 1))), link_space=4, rng=1),
              t.rand_itn(t.continuous_siteinds(t.named_binary_tree(4), map_dimension=2), link_space=3, rng=2)):
        assert _binarized(t.pack(f)) is None
    g = t.NamedGraph([(i, 1) for i in range(5)], [((0, 1), (i, 1)) for i in range(1, 5)])
    assert _binarized(t.pack(t.rand_itn(t.continuous_siteinds(g, map_dimension=2), link_space=9, rng=3))) is None
