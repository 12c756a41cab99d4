"""CPU tests of the host logic and of the C-ABI surface (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import itna_b200 as t
from itna_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ttneval.h")).read()
    declared = set(re.findall(r"\b(ttn_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_capi.EXPORTS)
    L = _capi.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.ttn_abi_version() == _capi.TTN_ABI_VERSION


def test_struct_layouts_match_header():
    # sizes computed by hand from include/ttneval.h (LP64)
    assert C.sizeof(_capi.ttn_desc) == 6 * 4 + 10 * 8
    assert C.sizeof(_capi.ttn_opts) == 4 * 4 + 8 + 16 + 4 + 4 + 4 + 4 + 8 + 4 + 4 + 8 + (4 + 4 + 8 + 4 + 4 + 8) + 16
    assert C.sizeof(_capi.ttn_grid) == 8 + 8 + 8 + 8 + 8
    assert C.sizeof(_capi.ttn_info) == 10 * 4 + 8 + 8 + 8


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    L = _capi.lib()
    if L.ttn_device_count() > 0:
        pytest.skip("GPU present")
    f = t.sin_itn(t.continuous_siteinds(t.named_grid((6, 1))))
    with pytest.raises(_capi.TTNError) as e:
        t.evaluate(f, 0.625)
    assert e.value.code == _capi.TTN_ERR_CUDA
    dfma, dmma = C.c_double(), C.c_double()
    assert L.ttn_measure_fp64_peak(0, C.byref(dfma), C.byref(dmma)) == _capi.TTN_ERR_CUDA


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "itensornumericalanalysis.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".jl", ".cpp")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.lower(), (dirpath, fn)


def test_graph_generators_vertex_order():
    assert t.named_grid((2, 3)).vertices() == [(1, 1), (2, 1), (1, 2), (2, 2), (1, 3), (2, 3)]
    g = t.named_comb_tree((2, 3))
    assert g.is_tree() and g.has_edge((1, 1), (2, 1)) and g.has_edge((2, 2), (2, 3))
    assert not g.has_edge((1, 2), (2, 2))
    b = t.named_binary_tree(7)
    assert b.nv() == 127 and b.is_tree()
    for seed in range(5):
        assert t.uniform_tree(12, rng=seed).is_tree()
    # default interleaving vs[i:map_dimension:L] (src/digit_inds.jl:17-21)
    s = t.continuous_siteinds(t.named_grid((6, 1)), map_dimension=2)
    assert s.dimension_vertices(1) == [(1, 1), (3, 1), (5, 1)]
    assert [s.vertex_digit(v) for v in s.dimension_vertices(2)] == [1, 2, 3]


def test_packer_layout_and_flop_rule():
    # BASELINE config 1: sin QTT, 20-bit MPS, chi=2 complex -> 592 flop/pt (SURVEY §8 d)
    s = t.continuous_siteinds(t.named_grid((20, 1)))
    p = t.pack(t.sin_itn(s))
    assert p.is_complex and p.flops_per_point() == 592
    assert sorted(p.link_dim.tolist()) == [1] + [2] * 19
    # it is rooted at an end: every vertex has at most one child
    assert np.bincount(p.parent[p.parent >= 0], minlength=20).max() == 1
    # BASELINE config 2: comb (2,30) with dimension i on tooth i, chi=16 -> 29 728
    g = t.named_comb_tree((2, 30))
    s = t.continuous_siteinds(g, [[(i, j) for j in range(1, 31)] for i in (1, 2)])
    p = t.pack(t.rand_itn(s, link_space=16, rng=0))
    assert p.flops_per_point() == 29728 and p.n_coords == 2
    assert np.bincount(p.parent[p.parent >= 0], minlength=60).max() == 1
    # BASELINE config 4: 28-site interleaved MPS, chi=32 -> 53 312
    s = t.continuous_siteinds(t.named_grid((28, 1)), map_dimension=2)
    p = t.pack(t.rand_itn(s, link_space=32, rng=0))
    assert p.flops_per_point() == 53312
    # thresholds are the reference's place values |v * b^-digit|
    for sidx, ind in enumerate(p.site_inds):
        k = s.digit(ind)
        assert p.thr[p.thr_ptr[sidx]:p.thr_ptr[sidx + 1]].tolist() == [0.0, 2.0 ** -k]
    # binary tree, 3-D: degree-3 vertices cost 2*chi^3 + 2*chi^2 (SURVEY §8 d)
    g = t.named_binary_tree(3)
    s = t.continuous_siteinds(g, [[v] for v in g.vertices()[1:4]] )
    chi = 4
    p = t.pack(t.rand_itn(s, link_space=chi, rng=0))
    assert p.flops_per_point() == 2 * (chi * chi + chi) + 2 * 2 * (chi ** 3 + chi ** 2)


def test_packer_errors():
    s = t.continuous_siteinds(t.named_grid((3, 3)))
    with pytest.raises(ValueError):          # loopy with chi > 1: not tree-contractible
        t.pack(t.const_itn(s, c=1.5, linkdim=4))
    t.pack(t.const_itn(s, c=1.5))            # loopy with chi = 1 is fine (links are trivial)
    s2 = t.continuous_siteinds(t.named_grid((6, 1)), map_dimension=2)
    f = t.rand_itn(s2, link_space=2, rng=0)
    with pytest.raises(KeyError):            # project() throws on a missing dimension (:90)
        t.pack(f, [1])
    with pytest.raises(AssertionError):      # length(xs) == length(dims), realindexmap.jl:68
        t.itensornetworkfunction._points_to_coords(f, [[0.1, 0.2, 0.3]], [1, 2])


def test_grid_points_matches_reference_formula():
    s = t.continuous_siteinds(t.named_grid((8, 1)), map_dimension=2)
    gp = s.grid_points(16, 1)      # L=4 bits per dimension: every index setting once
    assert gp == [i / 16 for i in range(16)]
    gp = s.grid_points(5, 1)       # a = round(16/5) = 3
    assert gp == [i * (3 / 16) for i in range(6)]


def test_julia_struct_mirrors_match_the_ctypes_layouts():
    """The Julia side (julia/TTNEvalB200.jl) is written, not executed here: keep its struct mirrors of ttn_desc /
    ttn_opts / ttn_grid in lock-step with the ctypes structs (which test_struct_layouts_match_header pins to the
    header) — same field order, same sizes — and its ccall symbols inside the exported set."""
    src = open(os.path.join(ROOT, "itensornumericalanalysis.jl_b200", "julia", "TTNEvalB200.jl")).read()
    size = {"Int32": 4, "Int64": 8, "Float32": 4, "Float64": 8}
    src = src.replace("mutable struct", "struct")

    def julia_fields(name):
        body = re.search(r"struct " + name + r"\n(.*?)\nend", src, re.S).group(1)
        out = []
        for line in body.splitlines():
            m = re.match(r"\s*(\w+)::([\w{}]+)", line)
            if m:
                out.append((m.group(1), 8 if m.group(2).startswith("Ptr") else size[m.group(2)]))
        return out

    def ctypes_fields(st):
        out = []
        for fname, ftype in st._fields_:
            n = C.sizeof(ftype)
            if issubclass(ftype, C.Array):      # sum_out[2] is two Float64 fields on the Julia side
                out += [(fname, C.sizeof(ftype._type_))] * ftype._length_
            else:
                out.append((fname, n))
        return out

    for jl, st in (("TTNDesc", _capi.ttn_desc), ("TTNOpts", _capi.ttn_opts), ("TTNGrid", _capi.ttn_grid)):
        a, b = julia_fields(jl), ctypes_fields(st)
        assert [s for _, s in a] == [s for _, s in b], (jl, a, b)
        for (ja, _), (cb, _) in zip(a, b):
            assert ja == cb or cb == "sum_out" and ja in ("sum_re", "sum_im"), (jl, ja, cb)
    called = set(re.findall(r"ccall\(\(:(ttn_\w+), LIBTTNEVAL\)", src))
    assert called <= set(_capi.EXPORTS)
    assert {"ttn_plan_create", "ttn_plan_destroy", "ttn_evaluate", "ttn_evaluate_grid", "ttn_evaluate_indices",
            "ttn_digits", "ttn_last_error"} <= called


def test_header_is_plain_c_and_the_c_example_links(tmp_path):
    """include/ttneval.h must be usable from any FFI: compile examples/c_abi_example.c as strict C11 against it,
    link libttneval.so and run it.  Without a GPU the example reports TTN_ERR_CUDA from ttn_plan_create (no CPU
    fallback) and exits 0; with one it prints f(x) = 1 + x_trunc."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    csrc = os.path.join(ROOT, "itensornumericalanalysis.jl_b200", "csrc")
    exe = str(tmp_path / "c_abi_example")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_example.c"), "-L" + csrc, "-lttneval",
                           "-Wl,-rpath," + csrc, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "libttneval ABI %d" % _capi.TTN_ABI_VERSION in r.stdout
    if _capi.lib().ttn_device_count() == 0:
        assert "no CUDA device" in r.stdout
    else:
        assert "f(0.375) = 1.375000" in r.stdout and "f(0.999) = 1.875000" in r.stdout


def test_plan_cache_is_dropped_when_a_vertex_is_assigned():
    """the reference mutates networks in place (psi[v] = ..., psi[v] *= c).  The plan cache is keyed
    on a version counter of the network, bumped by every vertex assignment — through the function object or through
    its .itensornetwork — so stale packed tensors are never evaluated."""
    import itna_b200 as t
    from itna_b200 import _capi
    s = t.continuous_siteinds(t.named_grid((6, 1)))
    f = t.rand_itn(s, link_space=3, rng=1)
    v = f.vertices()[0]
    for assign in (lambda: f.__setitem__(v, f[v] * 2.0), lambda: f.itensornetwork.__setitem__(v, f.itensornetwork[v] * 2.0)):
        f._plans[("stale",)] = object()
        f._plans_version = f.itensornetwork.version
        before = f.itensornetwork.version
        assign()
        assert f.itensornetwork.version != before
        try:
            f.plan()                      # no GPU here: plan creation itself fails, AFTER the stale entries are dropped
        except (_capi.TTNError, RuntimeError):
            pass
        assert ("stale",) not in f._plans
