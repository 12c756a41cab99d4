"""CPU tests of the integration mirror (src/integration.jl): partial_integrate gives the marginal of the dense
function over the full grid of the integrated dimensions; integrate is the all-dimensions mean / sum
(test/test_integration.jl:52-53 checks the same identity through evaluate)."""
import numpy as np
import pytest

import itna_b200 as t
import oracle as orc


def _grid(L):
    return np.arange(2 ** L) / 2.0 ** L


@pytest.mark.parametrize("kind", ["mps2d", "comb3", "complex"])
def test_partial_integrate_is_the_grid_marginal(kind):
    L = 4
    if kind == "mps2d":
        s = t.continuous_siteinds(t.named_grid((2 * L, 1)), map_dimension=2)
        f = t.rand_itn(s, link_space=3, rng=1, normalise=True)
        nd = 2
    elif kind == "comb3":
        g = t.named_comb_tree((3, L))
        s = t.continuous_siteinds(g, [[(j, i) for i in range(1, L + 1)] for j in range(1, 4)])
        f = t.rand_itn(s, link_space=2, rng=2, normalise=True) + t.cosh_itn(s, k=0.7, a=0.1, c=0.5, dim=3)
        nd = 3
    else:
        s = t.continuous_siteinds(t.named_grid((2 * L, 1)), map_dimension=2)
        f = t.sin_itn(s, k=1.3, a=0.2, c=0.9, dim=1) + t.cos_itn(s, k=0.4, a=0.0, c=1.1, dim=2)
        nd = 2
    xs = _grid(L)
    m = t.partial_integrate(f, [2])                       # mean over dimension 2
    assert m.indexmap.dimensions() == [d for d in range(1, nd + 1) if d != 2]
    rng = np.random.default_rng(0)
    keep = rng.random((12, nd - 1))
    pm = t.pack(m)
    got = orc.evaluate(pm, keep, orc.ORACLE_LD)
    pf = t.pack(f)
    for row, val in zip(keep, got):
        pts = np.empty((len(xs), nd))
        pts[:, 0] = row[0]
        pts[:, 1] = xs
        if nd == 3:
            pts[:, 2] = row[1]
        ref = orc.evaluate(pf, pts, orc.ORACLE_LD).mean()
        assert abs(val - ref) <= 1e-12 * max(abs(ref), 1e-3)
    # take_sum = true: the plain sum over the grid of dimension 2
    ms = t.partial_integrate(f, [2], take_sum=True)
    got_s = orc.evaluate(t.pack(ms), keep, orc.ORACLE_LD)
    assert np.allclose(got_s, got * 2 ** L, rtol=1e-13, atol=0)


def test_integrate_all_dimensions():
    L = 5
    s = t.continuous_siteinds(t.named_grid((2 * L, 1)), map_dimension=2)
    f = t.rand_itn(s, link_space=4, rng=3, normalise=True)
    xs = _grid(L)
    pts = np.array([[x, y] for x in xs for y in xs])
    vals = orc.evaluate(t.pack(f), pts, orc.ORACLE_LD)
    assert abs(t.integrate(f, take_sum=True) - vals.sum()) <= 1e-12 * np.abs(vals).sum()
    assert abs(t.integrate(f) - vals.mean()) <= 1e-12 * np.abs(vals).mean()
