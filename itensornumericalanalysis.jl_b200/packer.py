"""Host packer: ITensorNetworkFunction -> flat leaf-to-root description (`ttn_desc`).

Runs once per (network, dims).  It replaces what the reference redoes for EVERY point:
`copy(fitn)` (src/itensornetworkfunction.jl:85), the dictionary filters / sort of
`calculate_ind_values` (src/IndexMaps/realindexmap.jl:67-76) and the graph bookkeeping of
`project` (src/itensornetworkfunction.jl:84-94).  The Julia twin of this file is
julia/TTNEvalB200.jl (`pack`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .indexmaps import ComplexIndexMap


class PackedNetwork:
    """Flat arrays of include/ttneval.h:ttn_desc, kept alive together with the ctypes view."""

    def __init__(self, **kw):
        self.__dict__.update(kw)
        self._desc = None

    def desc(self):
        if self._desc is None:
            d = _capi.ttn_desc()
            d.abi_version = _capi.TTN_ABI_VERSION
            d.n_vertices = self.n_vertices
            d.n_coords = self.n_coords
            d.is_complex = int(self.is_complex)
            d.root = self.root
            d.n_sites = len(self.site_dim)
            i32, f64, i64 = C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int64)
            d.parent = self.parent.ctypes.data_as(i32)
            d.link_dim = self.link_dim.ctypes.data_as(i32)
            d.site_ptr = self.site_ptr.ctypes.data_as(i32)
            d.site_dim = self.site_dim.ctypes.data_as(i32)
            d.site_coord = self.site_coord.ctypes.data_as(i32)
            d.site_digit = self.site_digit.ctypes.data_as(i32)
            d.thr_ptr = self.thr_ptr.ctypes.data_as(i32)
            d.thr = self.thr.ctypes.data_as(f64)
            d.tensor_ptr = self.tensor_ptr.ctypes.data_as(i64)
            d.tensors = self.tensors.ctypes.data_as(C.c_void_p)
            self._desc = d
        return self._desc

    # SURVEY §8(d) flop rule: per non-leaf vertex, MACs = sum_j c_j*...*c_k*p.
    def flops_per_point(self):
        macs = 0
        children = {v: [] for v in range(self.n_vertices)}
        for v, p in enumerate(self.parent):
            if p >= 0:
                children[int(p)].append(v)
        for v in range(self.n_vertices):
            size = int(self.link_dim[v])
            for c in children[v]:
                size *= int(self.link_dim[c])
            for c in children[v]:
                macs += size
                size //= int(self.link_dim[c])
        return (8 if self.is_complex else 2) * macs


def _spanning_tree_edges(graph, linkdim):
    """Edges of a spanning tree that keeps every link of dimension > 1.  Any edge left out must
    have link dimension 1 (a trivial index: dropping it is exact).  Loopy networks with
    non-trivial loops are not tree-contractible and stay on the reference path
    (the reference itself only evaluates loopy graphs with chi=1 or alg="exact",
    test/test_realitensorfunction.jl:39-57)."""
    verts = graph.vertices()
    comp = {v: v for v in verts}

    def find(v):
        while comp[v] != v:
            comp[v] = comp[comp[v]]
            v = comp[v]
        return v

    kept, dropped = [], []
    edges = sorted(graph.edges(), key=lambda e: -linkdim(e))
    for a, b in edges:
        ra, rb = find(a), find(b)
        if ra == rb:
            if linkdim((a, b)) != 1:
                raise ValueError(
                    "network has a loop through links of dimension > 1: the batched evaluator "
                    "requires a tree (the reference asserts is_tree for tree algorithms, "
                    "src/itensornetworkfunction.jl:115)")
            dropped.append((a, b))
        else:
            comp[ra] = rb
            kept.append((a, b))
    if len(kept) != len(verts) - 1:
        raise ValueError("network graph is not connected")
    return kept, dropped


def _choose_root(verts, adj):
    """Path graphs (MPS): an end vertex, so the contraction is a single vector-matrix chain.
    Trees of maximum degree 3: the most central vertex of degree <= 2 (then no vertex has more than two
    children).  Other trees: a centre vertex (minimises the depth of the leaf-to-root schedule)."""
    if len(verts) == 1:
        return verts[0]
    deg = {v: len(adj[v]) for v in verts}
    if max(deg.values()) <= 2:
        ends = [v for v in verts if deg[v] == 1]
        return ends[-1]
    if max(deg.values()) == 3:
        # rooted at a vertex of degree <= 2, every vertex of such a tree has at most two children: the
        # per-vertex GEMM kernel applies (comb trees, binary trees).  Among those take the most central.
        def ecc(r):
            dist = {r: 0}
            todo = [r]
            for v in todo:
                for u in adj[v]:
                    if u not in dist:
                        dist[u] = dist[v] + 1
                        todo.append(u)
            return max(dist.values())
        cands = [v for v in verts if deg[v] <= 2]
        return min(cands, key=lambda v: (ecc(v), verts.index(v)))
    # peel leaves
    remaining = set(verts)
    d = dict(deg)
    leaves = [v for v in verts if d[v] == 1]
    while len(remaining) > 2:
        nxt = []
        for v in leaves:
            remaining.discard(v)
            for u in adj[v]:
                if u in remaining:
                    d[u] -= 1
                    if d[u] == 1:
                        nxt.append(u)
        leaves = nxt
    return [v for v in verts if v in remaining][0]


def pack(fitn, dims=None, root=None) -> PackedNetwork:
    """Pack `fitn` for points whose columns are the coordinates along `dims`
    (xs[i] is the coordinate along dims[i], src/IndexMaps/realindexmap.jl:67-71)."""
    tn = fitn.itensornetwork
    inm = fitn.indsnetworkmap
    imap = inm.indexmap
    if dims is None:
        dims = imap.dimensions()
    dims = [int(d) for d in dims]
    if len(set(dims)) != len(dims):
        raise ValueError("dims must not repeat a dimension")
    is_cmap = isinstance(imap, ComplexIndexMap)
    missing = [d for d in imap.dimensions() if d not in dims]
    if missing:
        # reference: project() throws a missing-key error (src/itensornetworkfunction.jl:90)
        raise KeyError(f"dims {dims} do not cover dimension(s) {missing} of the network")

    graph = tn.graph
    verts = graph.vertices()
    vid = {v: i for i, v in enumerate(verts)}
    kept, dropped = _spanning_tree_edges(graph, lambda e: tn.link(*e).dim)
    adj = {v: [] for v in verts}
    for a, b in kept:
        adj[a].append(b)
        adj[b].append(a)
    if root is None:
        root = _choose_root(verts, adj)
    n = len(verts)
    parent = np.full(n, -1, dtype=np.int32)
    order = [root]
    seen = {root}
    for v in order:
        for u in adj[v]:
            if u not in seen:
                seen.add(u)
                parent[vid[u]] = vid[v]
                order.append(u)
    dropped_links = {tn.link(a, b) for a, b in dropped}

    is_complex = tn.is_complex()
    dtype = np.complex128 if is_complex else np.float64
    link_dim = np.ones(n, dtype=np.int32)
    site_ptr = [0]
    site_dim, site_coord, site_digit, thr_ptr, thr = [], [], [], [0], []
    tensor_ptr = [0]
    blobs = []
    for v in verts:
        i = vid[v]
        t = tn[v]
        sites = list(inm[v])
        children = sorted((u for u in adj[v] if parent[vid[u]] == i), key=lambda u: vid[u])
        axes = sites + [tn.link(v, c) for c in children]
        if parent[i] >= 0:
            pl = tn.link(v, verts[parent[i]])
            axes.append(pl)
            link_dim[i] = pl.dim
        squeeze = [l for l in t.inds if l in dropped_links]
        arr = t.permute(axes + squeeze).array
        arr = arr.reshape([a.dim for a in axes])  # drop the trivial (dim 1) loop links
        blobs.append(np.ascontiguousarray(arr, dtype=dtype).reshape(-1))
        tensor_ptr.append(tensor_ptr[-1] + blobs[-1].size)
        for ind in sites:
            d = imap.dimension(ind)
            pos = dims.index(d)
            slot = 2 * pos + (0 if imap.is_real(ind) else 1) if is_cmap else pos
            site_dim.append(ind.dim)
            site_coord.append(slot)
            site_digit.append(imap.digit(ind))
            thr += [abs(imap.index_value_to_scalar(ind, k)) for k in range(ind.dim)]
            thr_ptr.append(len(thr))
        site_ptr.append(len(site_dim))

    return PackedNetwork(
        n_vertices=n,
        n_coords=(2 if is_cmap else 1) * len(dims),
        complex_coords=is_cmap,
        is_complex=is_complex,
        root=vid[root],
        dims=dims,
        vertex_names=verts,
        parent=parent,
        link_dim=link_dim,
        site_ptr=np.asarray(site_ptr, dtype=np.int32),
        site_dim=np.asarray(site_dim, dtype=np.int32),
        site_coord=np.asarray(site_coord, dtype=np.int32),
        site_digit=np.asarray(site_digit, dtype=np.int32),
        thr_ptr=np.asarray(thr_ptr, dtype=np.int32),
        thr=np.asarray(thr, dtype=np.float64),
        tensor_ptr=np.asarray(tensor_ptr, dtype=np.int64),
        tensors=np.concatenate(blobs) if blobs else np.zeros(0, dtype=dtype),
        site_inds=[ind for v in verts for ind in inm[v]],
    )
