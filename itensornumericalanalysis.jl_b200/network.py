"""Dense tensors with named indices and a tensor network on a NamedGraph.

A deliberately small stand-in for ITensors.ITensor / ITensorNetworks.ITensorNetwork: it holds
the data the evaluate path consumes (one dense array per vertex, site + link indices) and the
two pieces of network algebra the reference's function constructors need to build inputs for
that path: the direct sum `+` (ITensorNetworks `add`; algorithm documented by the reference's
un-included src/fixes.jl:39-91) and the vertex-wise product `multiply`
(src/elementary_operators.jl:260-281).  No contraction happens here: evaluation is done by the
CUDA library only.
"""
from __future__ import annotations

import numpy as np

from .graphs import NamedGraph
from .indexmaps import Index


class Tensor:
    __slots__ = ("array", "inds")

    def __init__(self, array, inds):
        array = np.asarray(array)
        inds = list(inds)
        assert array.ndim == len(inds), (array.shape, inds)
        assert tuple(i.dim for i in inds) == array.shape, (array.shape, inds)
        self.array = array
        self.inds = inds

    def permute(self, inds):
        inds = list(inds)
        assert set(inds) == set(self.inds) and len(inds) == len(self.inds)
        perm = [self.inds.index(i) for i in inds]
        return Tensor(np.transpose(self.array, perm), inds)

    def copy(self):
        return Tensor(self.array.copy(), list(self.inds))

    def __mul__(self, c):
        return Tensor(self.array * c, self.inds)

    __rmul__ = __mul__

    @property
    def dtype(self):
        return self.array.dtype


def outer(factors):
    """Outer product of (array, inds) factors -> Tensor over the concatenated indices."""
    arr = np.ones(())
    inds = []
    for a, ii in factors:
        arr = np.multiply.outer(arr, np.asarray(a))
        inds += list(ii)
    return Tensor(arr, inds)


def delta(inds, dtype=float):
    """delta(virt_inds): 1 where all index values coincide (src/utils.jl:42-47)."""
    inds = list(inds)
    if not inds:
        return Tensor(np.ones((), dtype=dtype), [])
    dims = [i.dim for i in inds]
    out = np.zeros(dims, dtype=dtype)
    for k in range(min(dims)):
        out[(k,) * len(dims)] = 1
    return Tensor(out, inds)


class TensorNetwork:
    """One tensor per vertex; every graph edge carries exactly one shared link Index."""

    def __init__(self, graph: NamedGraph, tensors: dict, links: dict):
        self.graph = graph
        self.tensors = tensors  # vertex -> Tensor
        self.links = links      # frozenset({a, b}) -> Index
        self.version = 0        # bumped by every vertex assignment; plan caches key on it

    def vertices(self):
        return self.graph.vertices()

    def link(self, a, b):
        return self.links[frozenset((a, b))]

    def __getitem__(self, v):
        return self.tensors[v]

    def __setitem__(self, v, t):
        self.tensors[v] = t
        self.version += 1

    def copy(self):
        return TensorNetwork(self.graph, {v: t.copy() for v, t in self.tensors.items()},
                             dict(self.links))

    def link_inds(self, v):
        return [self.link(v, u) for u in self.graph.neighbors(v)]

    def maxlinkdim(self):
        return max([i.dim for i in self.links.values()], default=1)

    def is_complex(self):
        return any(np.iscomplexobj(t.array) for t in self.tensors.values())


def make_links(graph, link_space):
    return {frozenset(e): Index(link_space, f"Link,{e[0]}-{e[1]}") for e in graph.edges()}


def random_tensornetwork(s, link_space=1, rng=None, dtype=float, normalise=False):
    """random_tensornetwork(rng, eltype, indsnetwork; link_space): i.i.d. N(0,1) entries
    (src/elementary_functions.jl:209-221 -> ITensorNetworks).  `normalise` scales each tensor
    by 1/sqrt(prod(link dims towards the leaves)) proxy = 1/sqrt(prod of all link dims / max)
    so that values stay O(1) through ~100 contractions (SURVEY §8 d)."""
    rng = np.random.default_rng(rng)
    graph = s.graph
    links = make_links(graph, link_space)
    tensors = {}
    for v in graph.vertices():
        inds = list(s[v]) + [links[frozenset((v, u))] for u in graph.neighbors(v)]
        shape = [i.dim for i in inds]
        arr = rng.standard_normal(shape)
        if np.issubdtype(np.dtype(dtype), np.complexfloating):
            arr = (arr + 1j * rng.standard_normal(shape)) / np.sqrt(2.0)
        if normalise:
            ld = [i.dim for i in inds[len(s[v]):]]
            if ld:
                arr = arr / np.sqrt(np.prod(ld) / max(ld))
        tensors[v] = Tensor(arr.astype(dtype), inds)
    return TensorNetwork(graph, tensors, links)


def add(tn1: TensorNetwork, tn2: TensorNetwork) -> TensorNetwork:
    """Direct sum on every link (block diagonal), identity on site indices: (tn1 + tn2)(x) =
    tn1(x) + tn2(x).  Algorithm as documented by src/fixes.jl:39-91."""
    assert tn1.graph.vertices() == tn2.graph.vertices()
    graph = tn1.graph
    assert set(map(frozenset, graph.edges())) == set(map(frozenset, tn2.graph.edges()))
    if graph.nv() == 1:
        (v,) = graph.vertices()
        t2 = tn2[v].permute(tn1[v].inds)
        return TensorNetwork(graph, {v: Tensor(tn1[v].array + t2.array, tn1[v].inds)}, {})
    new_links = {}
    for e in graph.edges():
        fe = frozenset(e)
        new_links[fe] = Index(tn1.links[fe].dim + tn2.links[fe].dim, f"Link,{e[0]}-{e[1]}")
    tensors = {}
    for v in graph.vertices():
        nbrs = graph.neighbors(v)
        l1 = [tn1.link(v, u) for u in nbrs]
        l2 = [tn2.link(v, u) for u in nbrs]
        sites = [i for i in tn1[v].inds if i not in l1]
        sites2 = [i for i in tn2[v].inds if i not in l2]
        assert set(sites) == set(sites2), "networks must share site indices"
        a1 = tn1[v].permute(sites + l1).array
        a2 = tn2[v].permute(sites + l2).array
        ln = [new_links[frozenset((v, u))] for u in nbrs]
        dtype = np.result_type(a1.dtype, a2.dtype)
        out = np.zeros([i.dim for i in sites] + [i.dim for i in ln], dtype=dtype)
        ns = len(sites)
        sl1 = (slice(None),) * ns + tuple(slice(0, i.dim) for i in l1)
        sl2 = (slice(None),) * ns + tuple(slice(i1.dim, i1.dim + i2.dim) for i1, i2 in zip(l1, l2))
        out[sl1] = a1
        out[sl2] = a2
        tensors[v] = Tensor(out, sites + ln)
    return TensorNetwork(graph, tensors, new_links)


def multiply(tn1: TensorNetwork, tn2: TensorNetwork) -> TensorNetwork:
    """Pointwise product of two functions on the same site indices: Hadamard on the shared site
    indices, Kronecker on the links (what operator_proj + contraction + combine_linkinds
    amounts to in src/elementary_operators.jl:247-281)."""
    assert tn1.graph.vertices() == tn2.graph.vertices()
    graph = tn1.graph
    new_links = {}
    for e in graph.edges():
        fe = frozenset(e)
        new_links[fe] = Index(tn1.links[fe].dim * tn2.links[fe].dim, f"Link,{e[0]}-{e[1]}")
    tensors = {}
    for v in graph.vertices():
        nbrs = graph.neighbors(v)
        l1 = [tn1.link(v, u) for u in nbrs]
        l2 = [tn2.link(v, u) for u in nbrs]
        sites = [i for i in tn1[v].inds if i not in l1]
        a1 = tn1[v].permute(sites + l1).array
        a2 = tn2[v].permute(sites + l2).array
        ns, nl = len(sites), len(nbrs)
        # out[s..., (l1_0,l2_0), (l1_1,l2_1), ...] = a1[s..., l1...] * a2[s..., l2...]
        e1 = a1.reshape(a1.shape[:ns] + tuple(x for i in l1 for x in (i.dim, 1)))
        e2 = a2.reshape(a2.shape[:ns] + tuple(x for i in l2 for x in (1, i.dim)))
        prod = e1 * e2
        ln = [new_links[frozenset((v, u))] for u in nbrs]
        out = prod.reshape([i.dim for i in sites] + [i.dim for i in ln])
        tensors[v] = Tensor(out, sites + ln)
        del nl
    return TensorNetwork(graph, tensors, new_links)
