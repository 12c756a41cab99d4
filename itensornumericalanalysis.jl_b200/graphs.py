"""Minimal named graphs — just enough of NamedGraphs.jl / Graphs.jl for the evaluate path.

The reference builds its networks on `named_grid`, `named_comb_tree`, `uniform_tree` and
`named_binary_tree` graphs (test/test_realitensorfunction.jl:6-7,66,129;
examples/2d_laplace_solver.jl:16-18).  Vertex ORDER matters: the default digit assignment
interleaves `vertices(g)` (src/digit_inds.jl:17-21), so the generators reproduce the vertex
order of NamedGraphs' generators (first tuple entry fastest, as Julia's CartesianIndices).
"""
from __future__ import annotations

import numpy as np


class NamedGraph:
    """Undirected graph with hashable vertex names and a fixed vertex order."""

    def __init__(self, vertices, edges=()):
        self._vertices = list(vertices)
        self._adj = {v: [] for v in self._vertices}
        if len(self._adj) != len(self._vertices):
            raise ValueError("duplicate vertex names")
        for a, b in edges:
            self.add_edge(a, b)

    # --- Graphs.jl-style queries -------------------------------------------------
    def vertices(self):
        return list(self._vertices)

    def nv(self):
        return len(self._vertices)

    def has_edge(self, a, b):
        return b in self._adj[a]

    def add_edge(self, a, b):
        if a == b:
            raise ValueError("self loops are not allowed")
        if not self.has_edge(a, b):
            self._adj[a].append(b)
            self._adj[b].append(a)

    def rem_edge(self, a, b):
        self._adj[a].remove(b)
        self._adj[b].remove(a)

    def neighbors(self, v):
        return list(self._adj[v])

    def edges(self):
        pos = {v: i for i, v in enumerate(self._vertices)}
        out = []
        for a in self._vertices:
            for b in self._adj[a]:
                if pos[a] < pos[b]:
                    out.append((a, b))
        return out

    def ne(self):
        return sum(len(n) for n in self._adj.values()) // 2

    def copy(self):
        return NamedGraph(self._vertices, self.edges())

    def is_connected(self):
        if not self._vertices:
            return True
        seen = {self._vertices[0]}
        stack = [self._vertices[0]]
        while stack:
            v = stack.pop()
            for u in self._adj[v]:
                if u not in seen:
                    seen.add(u)
                    stack.append(u)
        return len(seen) == len(self._vertices)

    def is_tree(self):
        return self.is_connected() and self.ne() == self.nv() - 1

    def rename_vertices(self, f):
        return NamedGraph([f(v) for v in self._vertices], [(f(a), f(b)) for a, b in self.edges()])

    def leaf_vertices(self):
        return [v for v in self._vertices if len(self._adj[v]) == 1]


def vertices(g):
    return g.vertices()


def is_tree(g):
    return g.is_tree()


def named_grid(dims):
    """named_grid((nx, ny)): vertices (i, j), i fastest; nearest-neighbour edges."""
    if isinstance(dims, int):
        dims = (dims,)
    dims = tuple(int(d) for d in dims)
    names = [tuple(int(x) + 1 for x in idx[::-1]) for idx in np.ndindex(*dims[::-1])]
    g = NamedGraph(names)
    for v in names:
        for ax in range(len(dims)):
            if v[ax] < dims[ax]:
                u = list(v)
                u[ax] += 1
                g.add_edge(v, tuple(u))
    return g


def named_comb_tree(dims):
    """named_comb_tree((nx, ny)): backbone (1,1)..(nx,1), tooth i = (i,1)..(i,ny)
    (consistent with examples/construct_multi_dimensional_function.jl:15-17)."""
    nx, ny = (int(d) for d in dims)
    names = [(i, j) for j in range(1, ny + 1) for i in range(1, nx + 1)]
    g = NamedGraph(names)
    for i in range(1, nx):
        g.add_edge((i, 1), (i + 1, 1))
    for i in range(1, nx + 1):
        for j in range(1, ny):
            g.add_edge((i, j), (i, j + 1))
    return g


def named_path_graph(n):
    g = NamedGraph(list(range(1, n + 1)))
    for i in range(1, n):
        g.add_edge(i, i + 1)
    return g


def named_binary_tree(depth):
    """Complete binary tree with `depth` levels (2^depth - 1 vertices); vertex names are the
    root-to-vertex paths (1,), (1,1), (1,2), (1,1,1), ... in breadth-first order."""
    names = [(1,)]
    edges = []
    frontier = [(1,)]
    for _ in range(depth - 1):
        nxt = []
        for v in frontier:
            for c in (1, 2):
                u = v + (c,)
                names.append(u)
                edges.append((v, u))
                nxt.append(u)
        frontier = nxt
    return NamedGraph(names, edges)


def uniform_tree(n, rng=None):
    """Random labelled tree on vertices 1..n from a uniformly random Pruefer sequence
    (what Graphs.uniform_tree does; test/test_realitensorfunction.jl:129)."""
    rng = np.random.default_rng(rng)
    g = NamedGraph(list(range(1, n + 1)))
    if n <= 1:
        return g
    if n == 2:
        g.add_edge(1, 2)
        return g
    code = [int(x) for x in rng.integers(1, n + 1, size=n - 2)]
    degree = {v: 1 for v in range(1, n + 1)}
    for c in code:
        degree[c] += 1
    for c in code:
        leaf = min(v for v in degree if degree[v] == 1)
        g.add_edge(leaf, c)
        degree[leaf] -= 1
        degree[c] -= 1
    u, w = [v for v in degree if degree[v] == 1]
    g.add_edge(u, w)
    return g
