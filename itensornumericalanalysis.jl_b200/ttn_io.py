"""Portable network files (`*.ttn.json`): one ITensorNetworkFunction — graph, site indices with their
index-map entries, vertex tensors — as plain JSON, so that the SAME network can be evaluated by this library
and by the Julia reference (`julia/ref_evaluate.jl` rebuilds an `ITensorNetworkFunction` from it with the
reference's own types and runs the reference's `evaluate`, src/itensornetworkfunction.jl:96-106).

Julia's `Random.seed!` stream cannot be reproduced outside Julia (SURVEY §8 d), so parity on random networks
has to travel in this direction: network written here, values produced there.  Numbers are written with
`repr` (shortest round-trip form): both JSON readers reproduce every double bit for bit.

Layout (format "ttn-json-1"):
  vertices  [[1, 1], [2, 1], ...]                       vertex names, in vertices(g) order
  edges     [[ia, ib, dim], ...]                        0-based vertex positions + link dimension
  map       "real" | "complex"
  sites     [{vertex, dim, dimension, digit, is_real}]  one entry per site index, grouped by vertex in the order
                                                        the indices sit on the vertex
  tensors   [{vertex, axes, shape, re, im?}]            axes: "s<k>" = site entry k, "e<k>" = edge k; data in
                                                        C order (last axis fastest) over `axes`
  dims      [1, 2, ...]                                 dimension numbers the coordinate columns refer to
  points    optional [[x_1, ..., x_D], ...]             (complex coordinates as [re, im] pairs)
"""
from __future__ import annotations

import json

import numpy as np

from .graphs import NamedGraph
from .indexmaps import ComplexIndexMap, Index, IndsNetwork, IndsNetworkMap, RealIndexMap
from .network import Tensor, TensorNetwork

FORMAT = "ttn-json-1"


def save_ttn(fitn, path, points=None, dims=None):
    tn, inm = fitn.itensornetwork, fitn.indsnetworkmap
    imap = inm.indexmap
    verts = list(tn.vertices())
    vpos = {v: i for i, v in enumerate(verts)}
    edges, edge_of = [], {}
    for key, ind in tn.links.items():
        a, b = sorted(key, key=lambda v: vpos[v])
        edge_of[ind] = len(edges)
        edges.append([vpos[a], vpos[b], ind.dim])
    is_cmap = isinstance(imap, ComplexIndexMap)
    sites, site_of = [], {}
    for v in verts:
        for ind in inm.indsnetwork[v]:
            site_of[ind] = len(sites)
            sites.append({"vertex": vpos[v], "dim": ind.dim, "dimension": int(imap.index_dimension[ind]),
                          "digit": int(imap.index_digit[ind]),
                          "is_real": bool(imap.index_real[ind]) if is_cmap else True})
    tensors = []
    for v in verts:
        t = tn[v]
        axes = [f"s{site_of[i]}" if i in site_of else f"e{edge_of[i]}" for i in t.inds]
        arr = np.ascontiguousarray(t.array)
        rec = {"vertex": vpos[v], "axes": axes, "shape": list(arr.shape),
               "re": [float(x) for x in arr.real.reshape(-1)]}
        if np.iscomplexobj(arr):
            rec["im"] = [float(x) for x in arr.imag.reshape(-1)]
        tensors.append(rec)
    doc = {"format": FORMAT, "vertices": [list(v) if isinstance(v, tuple) else [v] for v in verts], "edges": edges,
           "map": "complex" if is_cmap else "real", "sites": sites, "tensors": tensors,
           "dims": [int(d) for d in (dims if dims is not None else imap.dimensions())]}
    if points is not None:
        pts = np.asarray(points)
        if np.iscomplexobj(pts):
            doc["points"] = [[[float(z.real), float(z.imag)] for z in row] for row in pts]
        else:
            doc["points"] = [[float(x) for x in row] for row in pts]
    with open(path, "w") as fh:
        json.dump(doc, fh)
    return path


def load_ttn(path):
    """-> (ITensorNetworkFunction, dims, points or None).  Inverse of save_ttn (fresh Index identities)."""
    from .itensornetworkfunction import ITensorNetworkFunction
    with open(path) as fh:
        doc = json.load(fh)
    assert doc["format"] == FORMAT, doc.get("format")
    verts = [tuple(v) if len(v) > 1 else v[0] for v in doc["vertices"]]
    g = NamedGraph(verts, [(verts[a], verts[b]) for a, b, _ in doc["edges"]])
    link_inds = [Index(d, f"Link,e{k}") for k, (_, _, d) in enumerate(doc["edges"])]
    links = {frozenset((verts[a], verts[b])): link_inds[k] for k, (a, b, _) in enumerate(doc["edges"])}
    is_cmap = doc["map"] == "complex"
    site_inds, index_digit, index_dimension, index_real = [], {}, {}, {}
    per_vertex = {v: [] for v in verts}
    for k, s in enumerate(doc["sites"]):
        tag = "Digit" + ("" if not is_cmap else (",Real" if s["is_real"] else ",Imag"))
        ind = Index(s["dim"], tag)
        site_inds.append(ind)
        per_vertex[verts[s["vertex"]]].append(ind)
        index_digit[ind], index_dimension[ind], index_real[ind] = s["digit"], s["dimension"], s["is_real"]
    isn = IndsNetwork(g)
    for v in verts:
        isn[v] = per_vertex[v]
    imap = ComplexIndexMap(index_digit, index_dimension, index_real) if is_cmap else RealIndexMap(index_digit, index_dimension)
    tensors = {}
    for rec in doc["tensors"]:
        inds = [site_inds[int(a[1:])] if a[0] == "s" else link_inds[int(a[1:])] for a in rec["axes"]]
        arr = np.array(rec["re"], dtype=np.float64).reshape(rec["shape"])
        if "im" in rec:
            arr = arr + 1j * np.array(rec["im"], dtype=np.float64).reshape(rec["shape"])
        tensors[verts[rec["vertex"]]] = Tensor(arr, inds)
    f = ITensorNetworkFunction(TensorNetwork(g, tensors, links), IndsNetworkMap(isn, imap))
    pts = doc.get("points")
    if pts is not None:
        pts = np.array([[complex(*x) for x in row] for row in pts]) if is_cmap else np.array(pts, dtype=np.float64)
    return f, doc["dims"], pts
