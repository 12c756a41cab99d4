"""Function constructors that produce INPUTS for the evaluate path.

Network construction is out of scope for the kernels (SURVEY §2 marks it OOS: it runs once on
the host); these restatements exist so the parity tests can rebuild the reference's own
known-answer cases (test/test_realitensorfunction.jl, test/test_complexitensorfunction.jl)
without Julia.  Formulas follow src/elementary_functions.jl:16-135,209-261 and
src/utils.jl:42-47.
"""
from __future__ import annotations

import numpy as np

from .indexmaps import ComplexIndexMap, IndsNetworkMap
from .itensornetworkfunction import ITensorNetworkFunction
from .network import Tensor, TensorNetwork, delta, make_links, outer, random_tensornetwork


def default_c_value():
    return 1.0


def default_a_value():
    return 0.0


def default_k_value():
    return 1.0


def default_nterms():
    return 20


def default_dim():
    return 1


def c_tensor(phys_inds, virt_inds, dtype=float):
    """src/utils.jl:42-47 — delta over the virtual indices, ones over the physical ones."""
    assert len({i.dim for i in virt_inds}) <= 1
    ones = outer([(np.ones(i.dim), [i]) for i in phys_inds])
    d = delta(virt_inds, dtype=dtype)
    return Tensor(np.multiply.outer(ones.array, d.array), ones.inds + d.inds)


def const_itensornetwork(s: IndsNetworkMap, c=None, linkdim=1):
    """src/elementary_functions.jl:16-26 — f == c with every link of dimension `linkdim`."""
    c = default_c_value() if c is None else c
    graph = s.graph
    nv = graph.nv()
    links = make_links(graph, linkdim)
    cc = (complex(c) / linkdim) ** (1.0 / nv) if (np.iscomplexobj(c) or c < 0) else (c / linkdim) ** (1.0 / nv)
    tensors = {}
    for v in graph.vertices():
        virt = [links[frozenset((v, u))] for u in graph.neighbors(v)]
        tensors[v] = c_tensor(s[v], virt) * cc
    return ITensorNetworkFunction(TensorNetwork(graph, tensors, links), s)


def exp_itensornetwork(s: IndsNetworkMap, k=None, a=None, c=None, dim=None):
    """src/elementary_functions.jl:30-53 — product state for c*exp(k*x_dim + a)."""
    k = default_k_value() if k is None else k
    a = default_a_value() if a is None else a
    c = default_c_value() if c is None else c
    dim = default_dim() if dim is None else dim
    psi = const_itensornetwork(s)
    tn = psi.itensornetwork
    dim_vertices = s.dimension_vertices(dim)
    Lx = len(dim_vertices)
    for v in dim_vertices:
        sinds = s[v]
        linds = [i for i in tn[v].inds if i not in sinds]
        factors = []
        for sind in sinds:
            if s.dimension(sind) == dim:
                vals = np.exp(k * np.asarray(s.index_values_to_scalars(sind)))
                factors.append((vals, [sind]))
            else:
                factors.append((np.ones(sind.dim), [sind]))
        t = outer(factors)
        d = delta(linds)
        arr = np.multiply.outer(t.array, d.array) * np.exp(a / Lx)
        tn[v] = Tensor(arr, t.inds + d.inds)
    first = dim_vertices[0]
    tn[first] = tn[first] * c
    return psi


def cosh_itensornetwork(s, k=None, a=None, c=None, dim=None):
    k, a, c = _kac(k, a, c)
    return exp_itensornetwork(s, a=a, k=k, c=0.5 * c, dim=dim) + exp_itensornetwork(
        s, a=-a, k=-k, c=0.5 * c, dim=dim)


def sinh_itensornetwork(s, k=None, a=None, c=None, dim=None):
    k, a, c = _kac(k, a, c)
    return exp_itensornetwork(s, a=a, k=k, c=0.5 * c, dim=dim) + exp_itensornetwork(
        s, a=-a, k=-k, c=-0.5 * c, dim=dim)


def cos_itensornetwork(s, k=None, a=None, c=None, dim=None):
    k, a, c = _kac(k, a, c)
    return exp_itensornetwork(s, a=a * 1j, k=k * 1j, c=0.5 * c, dim=dim) + exp_itensornetwork(
        s, a=-a * 1j, k=-k * 1j, c=0.5 * c, dim=dim)


def sin_itensornetwork(s, k=None, a=None, c=None, dim=None):
    k, a, c = _kac(k, a, c)
    return exp_itensornetwork(s, a=a * 1j, k=k * 1j, c=-0.5j * c, dim=dim) + exp_itensornetwork(
        s, a=-a * 1j, k=-k * 1j, c=0.5j * c, dim=dim)


def tanh_itensornetwork(s, k=None, a=None, c=None, nterms=None, dim=None):
    """src/elementary_functions.jl:87-105 — 1 + sum_n 2(-1)^n exp(-2n(kx+a))."""
    k, a, c = _kac(k, a, c)
    nterms = default_nterms() if nterms is None else nterms
    dim = default_dim() if dim is None else dim
    psi = const_itensornetwork(s)
    first = s.dimension_vertices(dim)[0]
    for n in range(1, nterms + 1):
        t = exp_itensornetwork(s, a=-2 * n * a, k=-2 * k * n, dim=dim)
        t.itensornetwork[first] = t.itensornetwork[first] * (2 * ((-1) ** n))
        psi = psi + t
    psi.itensornetwork[first] = psi.itensornetwork[first] * c
    return psi


def _kac(k, a, c):
    return (default_k_value() if k is None else k, default_a_value() if a is None else a,
            default_c_value() if c is None else c)


def random_itensornetwork(s: IndsNetworkMap, link_space=1, rng=None, eltype=float,
                          normalise=False):
    """src/elementary_functions.jl:209-221."""
    return ITensorNetworkFunction(
        random_tensornetwork(s.indsnetwork, link_space=link_space, rng=rng, dtype=eltype,
                             normalise=normalise), s)


def delta_p(s: IndsNetworkMap, xs, dims=None):
    """src/elementary_functions.jl:224-261 — product state that is 1 on the digits of xs (planes
    when not all dimensions are given); a list of points gives the sum of their deltas."""
    if len(xs) and isinstance(xs[0], (list, tuple, np.ndarray)):
        points = xs
        points_dims = dims if dims is not None else [list(range(1, len(p) + 1)) for p in points]
        assert len(points) != 0 and len(points) == len(points_dims)
        out = None
        for p, d in zip(points, points_dims):
            t = delta_p(s, list(p), list(d))
            out = t if out is None else out + t
        return out
    if dims is None:
        dims = list(range(1, len(xs) + 1))
    ivmap = s.calculate_ind_values(list(xs), list(dims))
    graph = s.graph
    links = make_links(graph, 1)
    tensors = {}
    for v in graph.vertices():
        factors = []
        for sind in s[v]:
            if sind in ivmap:
                e = np.zeros(sind.dim)
                e[ivmap[sind]] = 1.0
            else:
                e = np.ones(sind.dim)
            factors.append((e, [sind]))
        t = outer(factors)
        virt = [links[frozenset((v, u))] for u in graph.neighbors(v)]
        d = delta(virt)
        tensors[v] = Tensor(np.multiply.outer(t.array, d.array), t.inds + d.inds)
    return ITensorNetworkFunction(TensorNetwork(graph, tensors, links), s)


const_itn = const_itensornetwork
exp_itn = exp_itensornetwork
cosh_itn = cosh_itensornetwork
sinh_itn = sinh_itensornetwork
tanh_itn = tanh_itensornetwork
cos_itn = cos_itensornetwork
sin_itn = sin_itensornetwork
rand_itn = random_itensornetwork
