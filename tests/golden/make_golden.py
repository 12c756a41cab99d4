"""Generates tests/golden/reference_known_answers.json.

The reference (Julia) cannot run in this image, so the golden vectors are the reference's OWN
known-answer tests restated with fixed parameters: every case below is a test of
/root/reference/test/test_realitensorfunction.jl, test_complexitensorfunction.jl or
test_indexmaps.jl (cited per case), with the `rand()` parameters replaced by seeded values and
the expected value computed analytically with mpmath at 50 digits (the reference asserts the same
analytic identity with rtol sqrt(eps)).  Nothing here reads /root/reference at run time.

    python tests/golden/make_golden.py
"""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 50
rng = np.random.default_rng(1234)  # the reference seeds Random.seed!(1234); streams differ, checks don't


def c2(z):
    z = complex(z)
    return [z.real, z.imag]


def mpc(z):
    if isinstance(z, (list, tuple)):
        return mp.mpc(z[0], z[1])
    return mp.mpc(*c2(z))


FUNCS = {"exp": mp.exp, "cosh": mp.cosh, "sinh": mp.sinh, "cos": mp.cos, "sin": mp.sin, "tanh": mp.tanh}


def trunc(x, base, L):
    """What the digit map does to x in [0,1): keep L base-`base` digits (floor)."""
    q = mp.floor(mp.mpf(x) * mp.mpf(base) ** L)
    return q / mp.mpf(base) ** L


def expected(terms, point, base, Ls, truncate=False):
    """sum_t c*f(k*z_dim + a).  The reference's tests assert the identity at the point itself
    (their points are representable in the available digits: 0.625, 5/9, ...), so by default z
    is used as given.  truncate=True (base 2 only, where the greedy loop is exactly floor):
    z is first cut to L digits, as evaluate does for non-dyadic points."""
    tot = mp.mpc(0)
    for t in terms:
        d = t["dim"] - 1
        z = point[d]
        if truncate:
            assert base == 2
            re = trunc(z[0], base, Ls[d][0])
            im = trunc(z[1], base, Ls[d][1]) if Ls[d][1] > 0 else mp.mpf(0)
        else:
            re, im = mp.mpf(z[0]), mp.mpf(z[1])
        zz = mp.mpc(re, im)
        if t["func"] == "const":
            tot += mpc(t["c"])
        else:
            tot += mpc(t["c"]) * FUNCS[t["func"]](mpc(t["k"]) * zz + mpc(t["a"]))
    return [float(tot.real), float(tot.imag)]


cases = []
names = ["cosh", "sinh", "exp", "cos", "sin"]

# --- test_realitensorfunction.jl:58-80 (binary) and :89-104 (trinary): comb tree (2,3), 1-D
for base, x, tag in ((2, 0.625, "binary"), (3, 5.0 / 9.0, "trinary")):
    for f in names:
        k, a, c = rng.random(3)
        terms = [dict(func=f, k=c2(k), a=c2(a), c=c2(c), dim=1)]
        cases.append(dict(name=f"real_{f}_{tag}", cite="test/test_realitensorfunction.jl:58-104",
                          graph=["comb", [2, 3]], map="real", base=base, map_dimension=1,
                          terms=terms, point=[[x, 0.0]], dims=[1],
                          expected=expected(terms, [[x, 0.0]], base, [(6, 0)])))
# --- :106-121 tanh, L=10 chain, nterms=50
k, a, c = rng.random(3)
terms = [dict(func="tanh", k=c2(k), a=c2(a), c=c2(c), dim=1, nterms=50)]
cases.append(dict(name="real_tanh", cite="test/test_realitensorfunction.jl:106-121",
                  graph=["grid", [10, 1]], map="real", base=2, map_dimension=1, terms=terms,
                  point=[[0.625, 0.0]], dims=[1], tol=1e-9,
                  expected=expected(terms, [[0.625, 0.0]], 2, [(10, 0)])))
# --- :39-50 const on a loopy 3x3 grid (chi = 1), and :149-161 3-D const
terms = [dict(func="const", c=c2(1.5), dim=1)]
cases.append(dict(name="real_const_loopy", cite="test/test_realitensorfunction.jl:39-50",
                  graph=["grid", [3, 3]], map="real", base=2, map_dimension=1, terms=terms,
                  point=[[0.5, 0.0]], dims=[1], expected=[1.5, 0.0]))
cases.append(dict(name="real_const_3d", cite="test/test_realitensorfunction.jl:149-161",
                  graph=["grid", [3, 3]], map="real", base=2, map_dimension=3, terms=terms,
                  point=[[0.5, 0.0], [0.25, 0.0], [0.0, 0.0]], dims=[1, 2, 3], expected=[1.5, 0.0]))
# --- :171-189 f(x) + f(y) on an interleaved L=10 chain at (0.625, 0.25)
for f in names:
    k, a, c = rng.random(3)
    terms = [dict(func=f, k=c2(k), a=c2(a), c=c2(c), dim=1), dict(func=f, k=c2(k), a=c2(a), c=c2(c), dim=2)]
    pt = [[0.625, 0.0], [0.25, 0.0]]
    cases.append(dict(name=f"real_{f}_2d_sum", cite="test/test_realitensorfunction.jl:171-189",
                      graph=["grid", [10, 1]], map="real", base=2, map_dimension=2, terms=terms,
                      point=pt, dims=[1, 2], expected=expected(terms, pt, 2, [(5, 0), (5, 0)])))
# --- non-dyadic points as test/test_operators.jl:35-45 uses (0.025, 0.1, 0.675): floor to L bits
for x in (0.025, 0.1, 0.675):
    k, a, c = rng.random(3)
    terms = [dict(func="sin", k=c2(k), a=c2(a), c=c2(c), dim=1)]
    cases.append(dict(name=f"real_sin_nondyadic_{x}", cite="test/test_operators.jl:35-45",
                      graph=["grid", [12, 1]], map="real", base=2, map_dimension=1, terms=terms,
                      point=[[x, 0.0]], dims=[1], expected=expected(terms, [[x, 0.0]], 2, [(12, 0)], truncate=True)))

# --- test_complexitensorfunction.jl:58-80 / :91-110: real and imaginary digits on different teeth
for base, z, tag in ((2, 0.625 + 0.25j, "binary"), (3, 5.0 / 9.0 + 4.0j / 9.0, "trinary")):
    for f in names:
        k, a, c = (rng.random(3) + 1j * rng.random(3))
        terms = [dict(func=f, k=c2(k), a=c2(a), c=c2(c), dim=1)]
        cases.append(dict(name=f"cplx_{f}_{tag}", cite="test/test_complexitensorfunction.jl:58-110",
                          graph=["comb", [2, 3]], map="complex", base=base,
                          real_dimension_vertices=[[[1, 1], [1, 2], [1, 3]]],
                          imag_dimension_vertices=[[[2, 1], [2, 2], [2, 3]]],
                          terms=terms, point=[c2(z)], dims=[1],
                          expected=expected(terms, [c2(z)], base, [(3, 3)])))
# --- :112-130 tanh with alternating real/imag vertices
k, a, c = (rng.random(3) + 1j * rng.random(3))
terms = [dict(func="tanh", k=c2(k), a=c2(a), c=c2(c), dim=1, nterms=50)]
cases.append(dict(name="cplx_tanh", cite="test/test_complexitensorfunction.jl:112-130",
                  graph=["grid", [10, 1]], map="complex", base=2,
                  real_dimension_vertices=[[[i, 1] for i in range(1, 11, 2)]],
                  imag_dimension_vertices=[[[i, 1] for i in range(2, 11, 2)]],
                  terms=terms, point=[c2(0.625 + 0.125j)], dims=[1], tol=1e-9,
                  expected=expected(terms, [c2(0.625 + 0.125j)], 2, [(5, 5)])))
# --- :183-207 two site indices per vertex (Real of dim 1 with Imag of dim 2), sum of two functions
L = 10
h = L // 2
rv = [[[i, 1] for i in range(1, h + 1)], [[i, 1] for i in range(h + 1, L + 1)]]
iv = [[[i, 1] for i in range(h + 1, L + 1)], [[i, 1] for i in range(1, h + 1)]]
for f in names:
    k, a, c = (rng.random(3) + 1j * rng.random(3))
    terms = [dict(func=f, k=c2(k), a=c2(a), c=c2(c), dim=1), dict(func=f, k=c2(k), a=c2(a), c=c2(c), dim=2)]
    pt = [c2(0.625 + 0.875j), c2(0.25 + 0.125j)]
    cases.append(dict(name=f"cplx_{f}_2d_sum", cite="test/test_complexitensorfunction.jl:183-207",
                      graph=["grid", [L, 1]], map="complex", base=2,
                      real_dimension_vertices=rv, imag_dimension_vertices=iv,
                      terms=terms, point=pt, dims=[1, 2],
                      expected=expected(terms, pt, 2, [(5, 5), (5, 5)])))
# --- :158-173 3-D const with a complex map
terms = [dict(func="const", c=c2(1.5), dim=1)]
cases.append(dict(name="cplx_const_3d", cite="test/test_complexitensorfunction.jl:158-173",
                  graph=["grid", [3, 3]], map="complex", base=2, map_dimension=3, terms=terms,
                  point=[c2(0.5 + 0.125j), c2(0.25 + 0.875j), c2(0.0)], dims=[1, 2, 3],
                  expected=[1.5, 0.0]))

# --- digit round trips pinned with `==` (test/test_indexmaps.jl:27-30, 44-47, 60-64)
digits = [
    dict(name="digits_real_4x4", cite="test/test_indexmaps.jl:18-30", graph=["grid", [4, 4]], map="real",
         map_dimension=1, point=[[0.625, 0.0]], dims=[1]),
    dict(name="digits_cplx_4x4", cite="test/test_indexmaps.jl:33-47", graph=["grid", [4, 4]], map="complex",
         map_dimension=1, point=[c2(0.625 + 0.5j)], dims=[1]),
    dict(name="digits_cplx_10d", cite="test/test_indexmaps.jl:50-64", graph=["grid", [10, 10]], map="complex",
         real_dimension_vertices=[[[i, j] for i in range(1, 11)] for j in range(1, 11)],
         imag_dimension_vertices=[[[i, j] for i in range(1, 11)] for j in range(1, 11)],
         point=[c2(0.5 + 0.125j), c2(0.75 + 0.625j)] + [c2(0.0)] * 8, dims=list(range(1, 11))),
]

out = dict(
    note="Known answers of the reference's own tests, restated with seeded parameters; see make_golden.py",
    value_cases=cases, digit_cases=digits)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_known_answers.json")
with open(path, "w") as fh:
    json.dump(out, fh, indent=1)
print(f"wrote {len(cases)} value cases and {len(digits)} digit cases to {path}")
