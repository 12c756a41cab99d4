"""Host-side mirror of src/integration.jl for the evaluate path: marginals of an ITensorNetworkFunction.

`partial_integrate(fitn, dims)` (src/integration.jl:35-54) contracts every site index of the dimensions `dims`
with the constant vector c (1/base: the mean over that digit; take_sum: 1) and returns the function of the
remaining dimensions — a network of the same shape whose integrated vertices simply carry fewer (or no) site
indices, which the packer and every kernel accept.  Evaluating the result at a batch of points gives the
marginal sums over the full grid of the integrated dimensions without touching that grid.
`integrate` (src/integration.jl:6-17) is the all-dimensions case; on the device it is also available as the
summed grid (`ttn_evaluate_grid`, identity tested in tests/test_gpu_parity.py)."""
from __future__ import annotations

import numpy as np

from .indexmaps import IndsNetworkMap
from .network import Tensor


def partial_integrate(fitn, dims, take_sum=False):
    from .itensornetworkfunction import ITensorNetworkFunction
    inm = fitn.indsnetworkmap
    dims = [int(d) for d in dims]
    c = 1.0 if take_sum else 1.0 / inm.base()
    tn = fitn.itensornetwork.copy()
    new_imap = inm.indexmap.copy()
    s_new = inm.indsnetwork.copy()
    for v in inm.vertices():
        gone = [i for i in inm.indsnetwork[v] if inm.indexmap.index_dimension[i] in dims]
        if not gone:
            continue
        t = tn[v]
        keep = [i for i in t.inds if i not in gone]
        axes = tuple(t.inds.index(i) for i in gone)
        tn[v] = Tensor(np.sum(t.array, axis=axes) * (c ** len(gone)), keep)   # fitn[v] *= ITensor(c, sind)
        for i in gone:
            new_imap = new_imap.rem_index(i)
        s_new[v] = [i for i in inm.indsnetwork[v] if i not in gone]
    return ITensorNetworkFunction(tn, IndsNetworkMap(s_new, new_imap))


def integrate(fitn, take_sum=False):
    """All dimensions: a scalar (contracted on the host in numpy by summing every site axis first)."""
    g = partial_integrate(fitn, fitn.indexmap.dimensions(), take_sum=take_sum)
    # no site indices are left: the value is the full contraction of the (tree) network, leaf by leaf
    tn = g.itensornetwork
    verts = list(tn.vertices())
    tensors = {v: tn[v] for v in verts}
    alive = set(verts)
    while len(alive) > 1:
        leaf = next(v for v in alive if sum(u in alive for u in tn.graph.neighbors(v)) == 1)
        (par,) = [u for u in tn.graph.neighbors(leaf) if u in alive]
        tl, tp = tensors[leaf], tensors[par]
        (link,) = [i for i in tl.inds if i in tp.inds]
        arr = np.tensordot(tp.array, tl.array, axes=([tp.inds.index(link)], [tl.inds.index(link)]))
        inds = [i for i in tp.inds if i != link] + [i for i in tl.inds if i != link]
        tensors[par] = Tensor(arr, inds)
        alive.remove(leaf)
    (last,) = alive
    return tensors[last].array.reshape(-1)[0] if tensors[last].array.size == 1 else tensors[last].array
