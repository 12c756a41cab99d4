"""CPU tests (no GPU): pin the oracle against the reference's own known answers and against two
independent restatements (host-side calculate_ind_values, dense numpy contraction)."""
import numpy as np
import pytest

import cases
import itna_b200 as t
import oracle as orc
from itna_b200 import _capi

GOLD = cases.load_golden()


@pytest.mark.parametrize("spec", GOLD["value_cases"], ids=lambda s: s["name"])
def test_oracle_matches_reference_known_answers(spec):
    f, point, dims, want = cases.build_golden_case(spec)
    packed = t.pack(f, dims)
    cmap = spec["map"] == "complex"
    pt = np.array([point])
    coords = np.empty((1, packed.n_coords))
    if cmap:
        coords[0, 0::2], coords[0, 1::2] = pt.real, pt.imag
    else:
        coords[0] = pt
    tol = spec.get("tol", 1e-12)
    for mode in (orc.ORACLE_LD, orc.ORACLE_F64, orc.ORACLE_BP):
        got = complex(orc.evaluate(packed, coords, mode)[0])
        assert abs(got - want) <= tol * max(1.0, abs(want)), (mode, got, want)


@pytest.mark.parametrize("spec", GOLD["digit_cases"], ids=lambda s: s["name"])
def test_digit_round_trip_exact(spec):
    """test/test_indexmaps.jl:27-30,44-47,60-64 — calculate_p(calculate_ind_values(x)) == x."""
    s = cases.build_siteinds(spec)
    cmap = spec["map"] == "complex"
    point = [complex(*z) if cmap else z[0] for z in spec["point"]]
    dims = spec["dims"]
    # host mirror
    m = s.calculate_ind_values(point, dims)
    assert set(m.keys()) == set(s.inds())
    back = s.calculate_p(m, dims)
    assert all(complex(b) == complex(p) for b, p in zip(back, point))  # exact ==
    # C oracle on the packed description (const network just to carry the index map)
    f = t.const_itn(s)
    packed = t.pack(f, dims)
    coords = np.empty((1, packed.n_coords))
    if cmap:
        coords[0, 0::2] = [complex(p).real for p in point]
        coords[0, 1::2] = [complex(p).imag for p in point]
    else:
        coords[0] = point
    dg = orc.digits(packed, coords)[0]
    for ind, d in zip(packed.site_inds, dg):
        assert m[ind] == int(d)
    # reconstruct from the threshold table: sum of chosen place values == coordinate, exactly
    rec = np.zeros(packed.n_coords)
    for sidx, d in enumerate(dg):
        rec[packed.site_coord[sidx]] += packed.thr[packed.thr_ptr[sidx] + d]
    assert (rec == coords[0]).all()


@pytest.mark.parametrize("case", cases.real_cases(), ids=lambda c: c[0])
def test_oracle_vs_host_digits_and_dense(case):
    name, f, dims, L = case
    rng = np.random.default_rng(5)
    pts = cases.edge_points(L, len(dims), rng, 40)
    packed = t.pack(f, dims)
    dg = orc.digits(packed, pts)
    for p, row in zip(pts[:25], dg[:25]):
        m = f.indsnetworkmap.calculate_ind_values(list(p), dims)
        assert [m[i] for i in packed.site_inds] == [int(x) for x in row]
    ref = orc.evaluate(packed, pts, orc.ORACLE_LD)
    if len(packed.site_dim) <= 16:  # dense brute force (cf. build_full_rank_tensor, src/utils.jl:28-39) for L <= 16
        dense = orc.dense_evaluate(f, pts, dims)
        assert orc.error_metric(dense, ref).max() < 1e-10
    assert orc.error_metric(orc.evaluate(packed, pts, orc.ORACLE_F64), ref).max() < 1e-12
    bp = orc.evaluate(packed, pts, orc.ORACLE_BP)
    assert orc.error_metric(bp, ref).max() < 1e-9  # BP's exp(sum log) is the less accurate order


@pytest.mark.parametrize("case", cases.complex_cases(), ids=lambda c: c[0])
def test_oracle_complex_vs_dense(case):
    name, f, dims, L = case
    rng = np.random.default_rng(6)
    z = cases.complex_points(L, len(dims), rng, 30)
    packed = t.pack(f, dims)
    coords = np.empty((z.shape[0], packed.n_coords))
    coords[:, 0::2], coords[:, 1::2] = z.real, z.imag
    ref = orc.evaluate(packed, coords, orc.ORACLE_LD)
    dense = orc.dense_evaluate(f, z, dims)
    assert orc.error_metric(dense, ref).max() < 1e-10
    dg = orc.digits(packed, coords)
    for p, row in zip(z[:20], dg[:20]):
        m = f.indsnetworkmap.calculate_ind_values(list(p), dims)
        assert [m[i] for i in packed.site_inds] == [int(x) for x in row]


def test_base2_digits_equal_bits_of_floor():
    """for base 2, contiguous digits and x in [0,1) the greedy loop gives bit k of
    floor(x * 2^L)."""
    L = 30
    s = t.continuous_siteinds(t.named_grid((L, 1)))
    packed = t.pack(t.const_itn(s))
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.random(20000), [0.0, 2.0 ** -L, 1 - 2.0 ** -L, 0.1, np.nextafter(0.5, 0),
                                            np.nextafter(1.0, 0)]])
    dg = orc.digits(packed, x[:, None])
    q = np.floor(x * 2.0 ** L).astype(np.int64)
    want = (q[:, None] >> (L - 1 - np.arange(L))[None, :]) & 1
    assert (dg == want).all()


def test_saturation_and_domain():
    s = t.continuous_siteinds(t.named_grid((8, 1)))
    packed = t.pack(t.const_itn(s))
    assert (orc.digits(packed, np.array([[1.0], [7.5]])) == 1).all()  # x >= 1 -> all ones
    for bad in (-0.25, np.nan):
        with pytest.raises(orc.OracleError) as e:
            orc.digits(packed, np.array([[bad]]))
        assert e.value.code == _capi.TTN_ERR_DOMAIN
    with pytest.raises(ValueError):
        s.calculate_ind_values(-0.25)


def test_delta_p_known_answers():
    """test/test_realitensorfunction.jl:210-258."""
    L = 10
    s = t.continuous_siteinds(t.named_grid((L, 1)), map_dimension=2)
    x0, y0 = 0.625, 0.25
    d = 2.0 ** -L
    xs = [0.0, d, 0.25, 0.5, 0.625, 0.875, 1 - d]

    def ev(f, pts):
        return orc.evaluate(t.pack(f, [1, 2]), np.array(pts, dtype=float), orc.ORACLE_LD)

    psi = t.delta_p(s, [x0, y0])
    assert ev(psi, [[x0, y0], [y0, x0]]).tolist() == [1.0, 0.0]
    psi = t.delta_p(s, [y0], [2])
    assert (ev(psi, [[x, y0] for x in xs]) == 1).all() and (ev(psi, [[x, 0.5] for x in xs]) == 0).all()
    psi = t.delta_p(s, [[x0, y0], [y0, x0]])
    assert ev(psi, [[x0, y0], [y0, x0], [0, 0], [0, y0]]).tolist() == [1.0, 1.0, 0.0, 0.0]
    psi = t.delta_p(s, [[x0, y0], [0.5]], [[1, 2], [2]])
    assert ev(psi, [[x0, y0]] + [[x, 0.5] for x in xs] + [[0, 0], [0, y0]]).tolist() == [1.0] * 8 + [0.0, 0.0]


def test_inv_pow_is_julias_pow_body():
    """float(b)^-k as Julia's `^(::Float64, ::Integer)` computes it (Base.Math.pow_body; realindexmap.jl:14,
    complexindexmap.jl:27) — the thresholds of the greedy loop.  n == -2 is the UNCOMPENSATED inv(x) * inv(x): one ulp
    above the correctly rounded value in bases 5 and 10 (constants from Julia >= 1.8), exact agreement in bases 3, 4, 6,
    7; powers of two are exact; every other exponent runs the compensated loop and is correctly rounded or one ulp off."""
    from fractions import Fraction
    from itna_b200.indexmaps import _inv_pow
    assert _inv_pow(5, 2) == 0.04000000000000001 and _inv_pow(10, 2) == 0.010000000000000002
    assert 0.04000000000000001 != 0.04 and 0.010000000000000002 != 0.01
    for b in (3, 4, 6, 7):
        assert _inv_pow(b, 2) == float(Fraction(1, b * b)), b
    for b in (2, 4, 8, 16):
        for k in range(0, 40):
            assert _inv_pow(b, k) == float(Fraction(1, b ** k))
    for b in (3, 5, 6, 7, 10):
        assert _inv_pow(b, 1) == 1.0 / b and _inv_pow(b, 0) == 1.0
        for k in range(3, 40):
            exact = Fraction(1, b ** k)
            got = _inv_pow(b, k)
            assert abs(Fraction(got) - exact) <= Fraction(float(exact)) * Fraction(1, 2 ** 52), (b, k)   # within one ulp
