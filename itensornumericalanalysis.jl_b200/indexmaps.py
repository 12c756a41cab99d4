"""Digit site indices, index maps and IndsNetworkMap — host-side mirror of the reference's
L1/L2 layers for the evaluate path.

Mirrors (same names, argument meaning and error behaviour):
  src/digit_inds.jl:17-21,60-115          default_dimension_vertices, digit_siteinds,
                                          complex_digit_siteinds
  src/IndexMaps/abstractindexmap.jl       dimensions, dimension, digit, dimension_inds,
                                          calculate_p, set_ind_values!, index_values_to_scalars
  src/IndexMaps/realindexmap.jl           RealIndexMap, index_value_to_scalar,
                                          calculate_ind_values, grid_points
  src/IndexMaps/complexindexmap.jl        ComplexIndexMap (Real/Imag tagged digits)
  src/indsnetworkmap.jl                   IndsNetworkMap, continuous_siteinds,
                                          complex_continuous_siteinds, vertex queries

`calculate_ind_values` here is the ONE-point host function the reference exposes (it is also
what `delta_p` uses to build networks).  The batched path never calls it: batched digits come
from the CUDA library (`ttn_digits` / fused into the contraction kernels).
"""
from __future__ import annotations

import itertools
from fractions import Fraction

import numpy as np

from .graphs import NamedGraph

_index_counter = itertools.count(1)


class Index:
    """An ITensors-style Index: a dimension, a tag string and a unique identity."""

    __slots__ = ("dim", "tags", "id")

    def __init__(self, dim, tags=""):
        self.dim = int(dim)
        self.tags = str(tags)
        self.id = next(_index_counter)

    def hastags(self, tag):
        return tag in [t.strip() for t in self.tags.split(",")]

    def __hash__(self):
        return hash(self.id)

    def __eq__(self, other):
        return isinstance(other, Index) and other.id == self.id

    def __repr__(self):
        return f"Index({self.dim}|{self.tags}|id={self.id})"


def dim(ind):
    return ind.dim


def hastags(ind, tag):
    return ind.hastags(tag)


class IndsNetwork:
    """Graph + site indices per vertex (ITensorNetworks.IndsNetwork, site part only)."""

    def __init__(self, graph: NamedGraph, site_space=None):
        self.graph = graph
        self.site_space = {v: [] for v in graph.vertices()}
        if site_space:
            for v, inds_v in site_space.items():
                self.site_space[v] = list(inds_v)

    def vertices(self):
        return self.graph.vertices()

    def edges(self):
        return self.graph.edges()

    def __getitem__(self, v):
        return self.site_space[v]

    def __setitem__(self, v, inds_v):
        self.site_space[v] = list(inds_v)

    def copy(self):
        return IndsNetwork(self.graph.copy(), {v: list(i) for v, i in self.site_space.items()})


def default_dimension_vertices(g, map_dimension=1):
    """src/digit_inds.jl:17-21 — vs[i:map_dimension:L] for i in 1:map_dimension."""
    vs = list(g.vertices())
    return [vs[i::map_dimension] for i in range(map_dimension)]


def _vertex_tag(v):
    return "×".join(str(x) for x in v) if isinstance(v, tuple) else str(v)


def digit_siteinds(g, dimension_vertices=None, base=2, map_dimension=1):
    """src/digit_inds.jl:72-86 — a vertex gets one Index per (dimension, digit) that names it."""
    if not dimension_vertices or len(dimension_vertices[0]) == 0:
        dimension_vertices = default_dimension_vertices(g, map_dimension=map_dimension)
    s = IndsNetwork(g)
    for d, verts in enumerate(dimension_vertices, start=1):
        for digit, v in enumerate(verts, start=1):
            s[v] = s[v] + [Index(base, f"Digit,V{_vertex_tag(v)},Dim{d},Dig{digit}")]
    return s


def complex_digit_siteinds(g, real_dimension_vertices=None, imag_dimension_vertices=None, base=2,
                           map_dimension=1):
    """src/digit_inds.jl:88-115 — Real-tagged indices first, then Imag-tagged ones."""
    if not real_dimension_vertices or len(real_dimension_vertices[0]) == 0:
        real_dimension_vertices = default_dimension_vertices(g, map_dimension=map_dimension)
    if not imag_dimension_vertices or len(imag_dimension_vertices[0]) == 0:
        imag_dimension_vertices = default_dimension_vertices(g, map_dimension=map_dimension)
    s = IndsNetwork(g)
    for d, verts in enumerate(real_dimension_vertices, start=1):
        for digit, v in enumerate(verts, start=1):
            s[v] = s[v] + [Index(base, f"Digit,Real,V{_vertex_tag(v)},Dim{d},Dig{digit}")]
    for d, verts in enumerate(imag_dimension_vertices, start=1):
        for digit, v in enumerate(verts, start=1):
            s[v] = s[v] + [Index(base, f"Digit,Imag,V{_vertex_tag(v)},Dim{d},Dig{digit}")]
    return s


def _fma(a, b, c):
    """Correctly rounded a*b + c (Python 3.12 has no math.fma): exact rational arithmetic, one rounding."""
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def _inv_pow(base, digit):
    """float(base)^-digit exactly as Julia computes it at realindexmap.jl:14 / complexindexmap.jl:27:
    `^(x::Float64, n::Integer)` -> Base.Math.pow_body (Julia >= 1.8; base/math.jl), restated operation by
    operation with fused muladd / two_mul (x86-64 with FMA, what Julia emits on every current host):
      n == -2 is the uncompensated inv(x)*inv(x)   (1 ulp off the correctly rounded value for bases 5 and 10),
      n == -1 and n <= -3 run the compensated square-and-multiply loop on inv(x).
    Powers of two are exact either way.  The Julia wrapper computes the thresholds with the reference's own
    index_value_to_scalar, so on the Julia side they are bit-identical by construction; this restatement makes
    the Python mirror (packer thresholds, test inputs) choose the same digits for every base."""
    x, n = float(base), -int(digit)
    if n == 0:
        return 1.0
    y, xnlo, ynlo = 1.0, 0.0, 0.0
    if n == 3:
        return x * x * x
    if n < 0:
        rx = 1.0 / x
        if n == -2:
            return rx * rx
        xnlo = -_fma(x, rx, -1.0) * rx
        x, n = rx, -n
    while n > 1:
        if n & 1:
            err = _fma(y, xnlo, x * ynlo)
            y, ynlo = x * y, _fma(x, y, -(x * y))
            ynlo += err
        err = x * 2 * xnlo
        x, xnlo = x * x, _fma(x, x, -(x * x))
        xnlo += err
        n >>= 1
    err = _fma(y, xnlo, x * ynlo)
    return _fma(x, y, err)


# --------------------------------------------------------------------------- index maps


class AbstractIndexMap:
    index_digit: dict
    index_dimension: dict

    def inds(self):
        assert self.index_dimension.keys() == self.index_digit.keys()
        return list(self.index_dimension.keys())

    def dimensions(self, inds=None):
        if inds is not None:
            return [self.index_dimension[i] for i in inds]
        out = []
        for d in self.index_dimension.values():  # unique, insertion order (abstractindexmap.jl:72)
            if d not in out:
                out.append(int(d))
        return out

    def dimension(self, ind=None):
        if ind is None:
            return max(self.dimensions())
        return self.index_dimension[ind]

    def digit(self, ind):
        return self.index_digit[ind]

    def digits(self, inds):
        return [self.index_digit[i] for i in inds]

    def index_values_to_scalars(self, ind):
        return [self.index_value_to_scalar(ind, i) for i in range(ind.dim)]

    def dimension_inds(self, dims):
        if isinstance(dims, (int, np.integer)):
            dims = [dims]
        return [i for i in self.index_dimension if self.index_dimension[i] in dims]

    def reduced_indexmap(self, dims):
        if isinstance(dims, (int, np.integer)):
            dims = [dims]
        keep = set(self.dimension_inds(dims))
        out = self.copy()
        for ind in self.inds():
            if ind not in keep:
                out = out.rem_index(ind)
        return out

    def calculate_p(self, ind_to_ind_value_map, dims=None):
        """abstractindexmap.jl:103-115 — inverse map, digits -> coordinate(s)."""
        single = isinstance(dims, (int, np.integer))
        if dims is None:
            dims = self.dimensions()
        if single:
            dims = [dims]
        out = []
        for d in dims:
            indices = [i for i in ind_to_ind_value_map if self.dimension(i) == d]
            out.append(sum(self.index_value_to_scalar(i, ind_to_ind_value_map[i]) for i in indices))
        return out

    def _set_ind_values(self, ind_to_ind_value_map, sorted_inds, x):
        """abstractindexmap.jl:121-138 — the greedy loop.  The reference never terminates for
        x<0 or NaN; we raise instead (documented deviation)."""
        x_rn = float(x)
        if not x_rn >= 0.0:
            raise ValueError(
                f"coordinate {x!r} is negative or NaN: the reference digit loop does not terminate")
        for ind in sorted_inds:
            ind_val = ind.dim - 1
            while True:
                thr = abs(self.index_value_to_scalar(ind, ind_val))
                if x_rn >= thr:
                    ind_to_ind_value_map[ind] = ind_val
                    x_rn -= thr
                    break
                ind_val -= 1

    def grid_points_all(self, d):
        """abstractindexmap.jl:149-155 — all base^L points of dimension d."""
        dims_ = [i.dim for i in self.dimension_inds(d)]
        assert all(y == dims_[0] for y in dims_)
        L = len(self.dimension_inds(d))
        return self.grid_points(dims_[0] ** L, d)


class RealIndexMap(AbstractIndexMap):
    """src/IndexMaps/realindexmap.jl:6-58."""

    def __init__(self, index_digit, index_dimension):
        self.index_digit = index_digit
        self.index_dimension = index_dimension

    @classmethod
    def from_dimension_indices(cls, dimension_indices):
        index_digit, index_dimension = {}, {}
        for d, indices in enumerate(dimension_indices, start=1):
            for bit, ind in enumerate(indices, start=1):
                index_digit[ind] = bit
                index_dimension[ind] = d
        return cls(index_digit, index_dimension)

    @classmethod
    def from_indsnetwork(cls, s, dimension_vertices=None):
        if dimension_vertices is None:
            dimension_vertices = default_dimension_vertices(s)
        dimension_indices = [[i for v in verts for i in s[v]] for verts in dimension_vertices]
        return cls.from_dimension_indices(dimension_indices)

    def index_value_to_scalar(self, ind, value):
        return value * _inv_pow(ind.dim, self.index_digit[ind])  # realindexmap.jl:13-15

    def copy(self):
        return RealIndexMap(dict(self.index_digit), dict(self.index_dimension))

    def ind(self, dim_, digit_):
        m = [i for i in self.index_dimension
             if self.index_dimension[i] == dim_ and self.index_digit[i] == digit_]
        (only,) = m
        return only

    def rem_index(self, ind):
        out = self.copy()
        del out.index_digit[ind]
        del out.index_dimension[ind]
        return out

    def merge(self, other):
        return RealIndexMap({**self.index_digit, **other.index_digit},
                            {**self.index_dimension, **other.index_dimension})

    def calculate_ind_values(self, xs, dims=None):
        """realindexmap.jl:67-76."""
        xs, dims = _normalise_xs_dims(self, xs, dims)
        assert len(xs) == len(dims)
        out = {}
        for d, x in zip(dims, xs):
            indices = self.dimension_inds(d)
            sorted_inds = sorted(indices, key=lambda i: self.index_digit[i])
            self._set_ind_values(out, sorted_inds, x)
        return out

    def grid_points(self, N=None, d=None):
        """realindexmap.jl:78-86."""
        if d is None:  # grid_points(imap, d)
            return self.grid_points_all(N)
        dims_ = [i.dim for i in self.dimension_inds(d)]
        assert all(y == dims_[0] for y in dims_)
        base = float(dims_[0])
        L = len(self.dimension_inds(d))
        a = float(np.round(base ** L / N))
        pts = [i * (a / base ** L) for i in range(0, N + 2)]
        return [x for x in pts if x < 1]

    def coordinate_slots(self, dims):
        """Real coordinate slots the packed description uses: one per entry of dims."""
        return len(dims)


class ComplexIndexMap(AbstractIndexMap):
    """src/IndexMaps/complexindexmap.jl:6-132."""

    def __init__(self, index_digit, index_dimension, index_real):
        self.index_digit = index_digit
        self.index_dimension = index_dimension
        self.index_real = index_real

    @classmethod
    def from_dimension_indices(cls, real_dimension_indices, imag_dimension_indices):
        index_digit, index_dimension, index_real = {}, {}, {}
        for d, indices in enumerate(real_dimension_indices, start=1):
            for bit, ind in enumerate(indices, start=1):
                index_digit[ind], index_dimension[ind], index_real[ind] = bit, d, True
        for d, indices in enumerate(imag_dimension_indices, start=1):
            for bit, ind in enumerate(indices, start=1):
                index_digit[ind], index_dimension[ind], index_real[ind] = bit, d, False
        return cls(index_digit, index_dimension, index_real)

    @classmethod
    def from_indsnetwork(cls, s, real_dimension_vertices=None, imag_dimension_vertices=None):
        if real_dimension_vertices is None:
            real_dimension_vertices = default_dimension_vertices(s)
        if imag_dimension_vertices is None:
            imag_dimension_vertices = default_dimension_vertices(s)
        re = [[i for v in verts for i in s[v] if i.hastags("Real")]
              for verts in real_dimension_vertices]
        im = [[i for v in verts for i in s[v] if i.hastags("Imag")]
              for verts in imag_dimension_vertices]
        return cls.from_dimension_indices(re, im)

    def is_real(self, ind):
        return self.index_real[ind]

    def real_indices(self, dim_=None):
        return [i for i in self.inds()
                if self.index_real[i] and (dim_ is None or self.index_dimension[i] == dim_)]

    def imaginary_indices(self, dim_=None):
        return [i for i in self.inds()
                if not self.index_real[i] and (dim_ is None or self.index_dimension[i] == dim_)]

    def index_value_to_scalar(self, ind, value):
        invb = _inv_pow(ind.dim, self.index_digit[ind])  # complexindexmap.jl:25-32
        return value * invb if self.index_real[ind] else 1j * (value * invb)

    def copy(self):
        return ComplexIndexMap(dict(self.index_digit), dict(self.index_dimension),
                               dict(self.index_real))

    def ind(self, dim_, digit_, real_ind=True):
        m = [i for i in self.inds()
             if self.index_dimension[i] == dim_ and self.index_digit[i] == digit_
             and self.index_real[i] == real_ind]
        (only,) = m
        return only

    def rem_index(self, ind):
        out = self.copy()
        del out.index_digit[ind]
        del out.index_dimension[ind]
        del out.index_real[ind]
        return out

    def merge(self, other):
        return ComplexIndexMap({**self.index_digit, **other.index_digit},
                               {**self.index_dimension, **other.index_dimension},
                               {**self.index_real, **other.index_real})

    def calculate_ind_values(self, xs, dims=None):
        """complexindexmap.jl:116-132 — real parts first, then imaginary parts."""
        xs, dims = _normalise_xs_dims(self, xs, dims)
        assert len(xs) == len(dims)
        out = {}
        for i, x in enumerate(xs):
            real_inds = sorted(self.real_indices(dims[i]), key=lambda j: self.index_digit[j])
            self._set_ind_values(out, real_inds, complex(x).real)
        for i, x in enumerate(xs):
            imag_inds = sorted(self.imaginary_indices(dims[i]), key=lambda j: self.index_digit[j])
            self._set_ind_values(out, imag_inds, complex(x).imag)
        return out

    def grid_points(self, N, d):
        """complexindexmap.jl:134-144 (the reference calls an undefined `imag_indices` there and
        throws; this is the evident intent)."""
        dims_ = [i.dim for i in self.dimension_inds(d)]
        assert all(y == dims_[0] for y in dims_)
        base = float(dims_[0])
        Lre, Lim = len(self.real_indices(d)), len(self.imaginary_indices(d))
        are, aim = float(np.round(base ** Lre / N)), float(np.round(base ** Lim / N))
        pts = [i * (are / base ** Lre) + 1j * j * (aim / base ** Lim)
               for i in range(0, N + 2) for j in range(0, N + 2)]
        return [z for z in pts if z.real < 1 and z.imag < 1]


def _normalise_xs_dims(imap, xs, dims):
    if isinstance(xs, (int, float, complex, np.number)):
        xs = [xs]
        if dims is None:
            dims = [imap.dimensions()[0]]  # first(dimensions(imap)), abstractindexmap.jl:140-143
        elif isinstance(dims, (int, np.integer)):
            dims = [dims]
    elif dims is None:
        dims = list(range(1, len(xs) + 1))  # abstractindexmap.jl:145-147
    return list(xs), [int(d) for d in dims]


# --------------------------------------------------------------------------- IndsNetworkMap


class IndsNetworkMap:
    """src/indsnetworkmap.jl:9-13 — {IndsNetwork, IndexMap} with the forwarded queries."""

    def __init__(self, indsnetwork: IndsNetwork, indexmap: AbstractIndexMap):
        self.indsnetwork = indsnetwork
        self.indexmap = indexmap

    # graph forwarding
    @property
    def graph(self):
        return self.indsnetwork.graph

    def vertices(self):
        return self.indsnetwork.vertices()

    def edges(self):
        return self.indsnetwork.edges()

    def nv(self):
        return self.graph.nv()

    def is_tree(self):
        return self.graph.is_tree()

    def __getitem__(self, v):
        return self.indsnetwork[v]

    def copy(self):
        return IndsNetworkMap(self.indsnetwork, self.indexmap)

    def inds(self, verts=None):
        """src/utils.jl:49-59."""
        if verts is None:
            verts = self.vertices()
        elif not isinstance(verts, list):
            return list(self.indsnetwork[verts])
        return [i for v in verts for i in self.indsnetwork[v]]

    def base(self):
        dims_ = [i.dim for i in self.inds()]
        assert all(d == dims_[0] for d in dims_)
        return dims_[0]

    def indexmaptype(self):
        return type(self.indexmap)

    # index-map forwarding (indsnetworkmap.jl:87-106)
    def ind(self, *a, **k):
        return self.indexmap.ind(*a, **k)

    def dimension(self, *a):
        return self.indexmap.dimension(*a)

    def dimensions(self, *a):
        return self.indexmap.dimensions(*a)

    def dimension_inds(self, dims):
        return self.indexmap.dimension_inds(dims)

    def digit(self, ind):
        return self.indexmap.digit(ind)

    def digits(self, inds):
        return self.indexmap.digits(inds)

    def calculate_ind_values(self, xs, dims=None):
        return self.indexmap.calculate_ind_values(xs, dims)

    def calculate_p(self, m, dims=None):
        return self.indexmap.calculate_p(m, dims)

    def grid_points(self, *a):
        return self.indexmap.grid_points(*a)

    def index_value_to_scalar(self, ind, value):
        return self.indexmap.index_value_to_scalar(ind, value)

    def index_values_to_scalars(self, ind):
        return self.indexmap.index_values_to_scalars(ind)

    # vertex queries (indsnetworkmap.jl:115-150)
    def vertices_dimensions(self, verts):
        return [self.dimension(i) for i in self.inds(list(verts))]

    def vertices_digits(self, verts):
        return [self.digit(i) for i in self.inds(list(verts))]

    def vertex_dimensions(self, v):
        return [self.dimension(i) for i in self.indsnetwork[v]]

    def vertex_digits(self, v):
        return [self.digit(i) for i in self.indsnetwork[v]]

    def vertex_dimension(self, v):
        (only,) = self.indsnetwork[v]
        return self.dimension(only)

    def vertex_digit(self, v):
        (only,) = self.indsnetwork[v]
        return self.digit(only)

    def dimension_vertices(self, dimension):
        if isinstance(dimension, (int, np.integer)):
            return [v for v in self.vertices() if dimension in self.vertex_dimensions(v)]
        return [v for v in self.vertices() if self.vertex_dimension(v) in dimension]

    def vertex(self, dimension, digit):
        index = self.ind(dimension, digit)
        (only,) = [v for v in self.vertices() if index in self.indsnetwork[v]]
        return only

    def reduced_indsnetworkmap(self, dims):
        """indsnetworkmap.jl:29-42 — keep only the indices (and vertices) of `dims`."""
        im = self.indexmap.reduced_indexmap(dims)
        keep = set(im.inds())
        g = self.graph.copy()
        site_space = {}
        verts = []
        for v in self.vertices():
            c = [i for i in self.indsnetwork[v] if i in keep]
            if c:
                site_space[v] = c
                verts.append(v)
        sub = NamedGraph(verts, [(a, b) for a, b in g.edges() if a in site_space and b in site_space])
        return IndsNetworkMap(IndsNetwork(sub, site_space), im)

    def rename_vertices(self, f):
        s = IndsNetwork(self.graph.rename_vertices(f),
                        {f(v): self.indsnetwork[v] for v in self.vertices()})
        return IndsNetworkMap(s, self.indexmap)


def RealIndsNetworkMap(g_or_s, dimension_vertices=None, base=2, map_dimension=1):
    """indsnetworkmap.jl:53-64 (aliases continuous_siteinds / real_continuous_siteinds)."""
    if isinstance(g_or_s, IndsNetwork):
        s = g_or_s
        if dimension_vertices is None:
            dimension_vertices = default_dimension_vertices(s, map_dimension=map_dimension)
    else:
        s = digit_siteinds(g_or_s, dimension_vertices, base=base, map_dimension=map_dimension)
        if dimension_vertices is None:
            dimension_vertices = default_dimension_vertices(g_or_s, map_dimension=map_dimension)
    return IndsNetworkMap(s, RealIndexMap.from_indsnetwork(s, dimension_vertices))


def ComplexIndsNetworkMap(g_or_s, real_dimension_vertices=None, imag_dimension_vertices=None,
                          base=2, map_dimension=1):
    """indsnetworkmap.jl:66-81 (alias complex_continuous_siteinds)."""
    if isinstance(g_or_s, IndsNetwork):
        s = g_or_s
        graph = s.graph
    else:
        graph = g_or_s
        s = complex_digit_siteinds(graph, real_dimension_vertices, imag_dimension_vertices,
                                   base=base, map_dimension=map_dimension)
    if real_dimension_vertices is None:
        real_dimension_vertices = default_dimension_vertices(graph, map_dimension=map_dimension)
    if imag_dimension_vertices is None:
        imag_dimension_vertices = default_dimension_vertices(graph, map_dimension=map_dimension)
    return IndsNetworkMap(
        s, ComplexIndexMap.from_indsnetwork(s, real_dimension_vertices, imag_dimension_vertices))


continuous_siteinds = RealIndsNetworkMap
real_continuous_siteinds = RealIndsNetworkMap
complex_continuous_siteinds = ComplexIndsNetworkMap
