"""CPU test of the portable network file (itna_b200.save_ttn / load_ttn): what travels to the Julia reference
(itensornumericalanalysis.jl_b200/julia/ref_evaluate.jl) is exactly the network this library evaluates."""
import json

import numpy as np
import pytest

import cases
import itna_b200 as t
import oracle as orc

CASES = {c[0]: (c, False) for c in cases.real_cases()}
CASES.update({c[0]: (c, True) for c in cases.complex_cases()})


@pytest.mark.parametrize("name", ["mps2d_chi8", "comb3x4_chi4", "bintree4_chi5", "base3_comb", "sin_qtt20", "sum_chi3p2",
                                  "cplx_teeth", "cplx_2site", "cplx_comb3x3", "single_vertex"])
def test_round_trip_preserves_digits_and_values(name, tmp_path):
    (_, f, dims, L), cplx = CASES[name]
    rng = np.random.default_rng(2)
    pts = cases.complex_points(L, len(dims), rng, 40) if cplx else cases.edge_points(L, len(dims), rng, 40)
    path = t.save_ttn(f, str(tmp_path / "net.ttn.json"), points=pts, dims=dims)
    g, dims2, pts2 = t.load_ttn(path)
    assert dims2 == list(dims) and (pts2 == pts).all()          # doubles survive bit for bit
    pa, pb = t.pack(f, dims), t.pack(g, dims2)
    assert pa.n_vertices == pb.n_vertices and pa.is_complex == pb.is_complex and pa.n_coords == pb.n_coords
    if pa.complex_coords:
        z = pts.astype(np.complex128)
        coords = np.empty((z.shape[0], 2 * z.shape[1]))
        coords[:, 0::2], coords[:, 1::2] = z.real, z.imag
    else:
        coords = pts.astype(np.float64)
    va, vb = orc.evaluate(pa, coords, orc.ORACLE_LD), orc.evaluate(pb, coords, orc.ORACLE_LD)
    assert orc.error_metric(vb, va).max() < 1e-13
    # the same tensors, element for element (the site-index order on a vertex and the edge set survive)
    doc = json.load(open(path))
    assert doc["format"] == "ttn-json-1" and len(doc["tensors"]) == pa.n_vertices
    total = sum(len(r["re"]) for r in doc["tensors"])
    assert total == sum(tt.array.size for tt in f.itensornetwork.tensors.values())
