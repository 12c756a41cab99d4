"""The reference's own test files, restated line by line against the Python mirror + CUDA path.

Each test below is the testset of the same name in /root/reference/test/test_realitensorfunction.jl
or test_complexitensorfunction.jl (cited), with `rand()` replaced by a seeded generator and the
reference's `≈` (rtol sqrt(eps) ~ 1.5e-8) tightened to 1e-12.  `evaluate` is the batched CUDA path
(one point per call here, exactly like the reference's tests)."""
import numpy as np
import pytest

import itna_b200 as t
from itna_b200 import evaluate

pytestmark = pytest.mark.gpu
rng = np.random.default_rng(1234)  # Random.seed!(1234) in the reference


def rand():
    return float(rng.random())


def crand():
    return rand() + 1j * rand()


def approx(a, b, tol=1e-12):
    return abs(complex(a) - complex(b)) <= tol * max(1.0, abs(complex(b)))


FUNCS = [("cosh", t.cosh_itn, np.cosh), ("sinh", t.sinh_itn, np.sinh), ("exp", t.exp_itn, np.exp),
         ("cos", t.cos_itn, np.cos), ("sin", t.sin_itn, np.sin)]


# ---- test/test_realitensorfunction.jl -------------------------------------------------------

def test_real_const():
    """:39-57 — const on a (loopy, chi = 1) 3x3 grid."""
    s = t.continuous_siteinds(t.named_grid((3, 3)))
    c = 1.5
    assert approx(evaluate(t.const_itn(s, c=c), 0.5), c)
    with pytest.raises(ValueError):  # linkdim = 4 on a loopy graph needs alg="exact": stays on the reference path
        evaluate(t.const_itn(s, c=c, linkdim=4), 0.5)


@pytest.mark.parametrize("name,net_func,func", FUNCS)
def test_real_elementary_binary(name, net_func, func):
    """:58-80 — comb tree (2,3), x = 0.625."""
    s = t.continuous_siteinds(t.named_comb_tree((2, 3)))
    a, k, c = rand(), rand(), rand()
    x = 0.625
    assert approx(evaluate(net_func(s, k=k, a=a, c=c), x), c * func(k * x + a))


@pytest.mark.parametrize("name,net_func,func", FUNCS)
def test_real_elementary_trinary(name, net_func, func):
    """:82-104 — base 3, x = 5/9."""
    s = t.continuous_siteinds(t.named_comb_tree((2, 3)), base=3)
    a, k, c = rand(), rand(), rand()
    x = 5.0 / 9.0
    assert approx(evaluate(net_func(s, k=k, a=a, c=c), x), c * func(k * x + a))


def test_real_tanh():
    """:106-121 — nterms = 50 on an L = 10 chain (series truncation allows 1e-9)."""
    s = t.continuous_siteinds(t.named_grid((10, 1)))
    a, k, c = rand(), rand(), rand()
    x = 0.625
    assert approx(evaluate(t.tanh_itn(s, k=k, a=a, c=c, nterms=50), x), c * np.tanh(k * x + a), 1e-9)


def test_real_const_3d():
    """:149-161."""
    s = t.continuous_siteinds(t.named_grid((3, 3)), map_dimension=3)
    assert approx(evaluate(t.const_itn(s, c=1.5), [0.5, 0.25, 0.0], [1, 2, 3]), 1.5)


@pytest.mark.parametrize("name,net_func,func", FUNCS)
def test_real_2d_sum(name, net_func, func):
    """:163-189 — f(x) + f(y) on an interleaved L = 10 chain."""
    s = t.continuous_siteinds(t.named_grid((10, 1)), map_dimension=2)
    x, y = 0.625, 0.25
    a, k, c = rand(), rand(), rand()
    psi = net_func(s, k=k, a=a, c=c, dim=1) + net_func(s, k=k, a=a, c=c, dim=2)
    assert approx(evaluate(psi, [x, y], [1, 2]), c * func(k * x + a) + c * func(k * y + a))


def test_real_tanh_2d():
    """:191-208 — named_grid((3,2)): a 2x3 ladder is loopy; with chi > 1 it needs alg="exact" in the
    reference, so here it must be rejected (documented: loopy networks stay on the reference path)."""
    s = t.continuous_siteinds(t.named_grid((3, 2)), map_dimension=2)
    psi = t.tanh_itn(s, k=rand(), a=rand(), c=rand(), nterms=20, dim=1)
    with pytest.raises(ValueError):
        evaluate(psi, [0.625, 0.875], [1, 2])


def test_real_delta_p():
    """:210-258."""
    L = 10
    s = t.continuous_siteinds(t.named_grid((L, 1)), map_dimension=2)
    x0, y0 = 0.625, 0.25
    delta = 2.0 ** (-1.0 * L)
    xs = [0.0, delta, 0.25, 0.5, 0.625, 0.875, 1 - delta]
    psi = t.delta_p(s, [x0, y0])
    assert approx(evaluate(psi, [x0, y0], [1, 2]), 1) and approx(evaluate(psi, [y0, x0], [1, 2]), 0)
    psi = t.delta_p(s, [y0], [2])
    assert all(approx(evaluate(psi, [x, y0], [1, 2]), 1) for x in xs)
    assert all(approx(evaluate(psi, [x, 0.5], [1, 2]), 0) for x in xs)
    psi = t.delta_p(s, [[x0, y0], [y0, x0]])
    assert approx(evaluate(psi, [x0, y0], [1, 2]), 1) and approx(evaluate(psi, [y0, x0], [1, 2]), 1)
    assert approx(evaluate(psi, [0, 0], [1, 2]), 0) and approx(evaluate(psi, [0, y0], [1, 2]), 0)
    p0 = 0.5
    psi = t.delta_p(s, [[x0, y0], [p0]], [[1, 2], [2]])
    assert approx(evaluate(psi, [x0, y0], [1, 2]), 1)
    assert all(approx(evaluate(psi, [x, p0], [1, 2]), 1) for x in xs)
    assert approx(evaluate(psi, [0, 0], [1, 2]), 0) and approx(evaluate(psi, [0, y0], [1, 2]), 0)


# ---- test/test_complexitensorfunction.jl ----------------------------------------------------

def test_complex_const():
    """:38-56."""
    s = t.complex_continuous_siteinds(t.named_grid((3, 3)))
    assert approx(evaluate(t.const_itn(s, c=1.5), 0.5 + 0.625j), 1.5)


@pytest.mark.parametrize("base,z", [(2, 0.625 + 0.25j), (3, 5.0 / 9.0 + 4.0j / 9.0)])
@pytest.mark.parametrize("name,net_func,func", FUNCS)
def test_complex_elementary(name, net_func, func, base, z):
    """:57-110 — real digits on one tooth, imaginary digits on the other."""
    g = t.named_comb_tree((2, 3))
    s = t.complex_continuous_siteinds(g, [[(1, 1), (1, 2), (1, 3)]], [[(2, 1), (2, 2), (2, 3)]], base=base)
    a, k, c = crand(), crand(), crand()
    assert approx(evaluate(net_func(s, k=k, a=a, c=c), z), c * func(k * z + a))


def test_complex_tanh():
    """:112-130 — alternating real / imaginary vertices."""
    L = 10
    s = t.complex_continuous_siteinds(t.named_grid((L, 1)), [[(i, 1) for i in range(1, L + 1, 2)]],
                                      [[(i, 1) for i in range(2, L + 1, 2)]])
    a, k, c = crand(), crand(), crand()
    z = 0.625 + 0.125j
    assert approx(evaluate(t.tanh_itn(s, k=k, a=a, c=c, nterms=50), z), c * np.tanh(k * z + a), 1e-8)


def test_complex_const_3d():
    """:158-173."""
    s = t.complex_continuous_siteinds(t.named_grid((3, 3)), map_dimension=3)
    z = [0.5 + 0.125j, 0.25 + 0.875j, 0.0]
    assert approx(evaluate(t.const_itn(s, c=1.5), z, [1, 2, 3]), 1.5)


@pytest.mark.parametrize("name,net_func,func", FUNCS)
def test_complex_2d_sum_two_site_indices_per_vertex(name, net_func, func):
    """:175-207 — each vertex carries a Real index of one dimension and an Imag index of the other."""
    L = 10
    h = L // 2
    rv = [[(i, 1) for i in range(1, h + 1)], [(i, 1) for i in range(h + 1, L + 1)]]
    iv = [[(i, 1) for i in range(h + 1, L + 1)], [(i, 1) for i in range(1, h + 1)]]
    s = t.complex_continuous_siteinds(t.named_grid((L, 1)), rv, iv)
    z1, z2 = 0.625 + 0.875j, 0.25 + 0.125j
    a, k, c = crand(), crand(), crand()
    psi = net_func(s, k=k, a=a, c=c, dim=1) + net_func(s, k=k, a=a, c=c, dim=2)
    assert approx(evaluate(psi, [z1, z2], [1, 2]), c * func(k * z1 + a) + c * func(k * z2 + a))

