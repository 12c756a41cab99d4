"""ITensorNetworkFunction and `evaluate` — host-side mirror of
src/itensornetworkfunction.jl for the batched path.

`evaluate(fitn, xs, dims)` keeps the reference's call forms
(src/itensornetworkfunction.jl:96-112):

    evaluate(fitn, x)                    one 1-D point            -> scalar
    evaluate(fitn, [x, y], [1, 2])       one point, xs[i] along dims[i] -> scalar
    evaluate(fitn, points, dims)         points: (npts, D) array or list of points -> ndarray

The batched form is the new method SURVEY §8(b) specifies (`Vector{<:Vector}` / matrix of
points; precedent `delta_p(s, points::Vector{<:Vector}, ...)`,
src/elementary_functions.jl:249-261).  All three run on the GPU through libttneval.so; there is
no CPU implementation behind them.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .indexmaps import ComplexIndexMap, IndsNetworkMap, RealIndsNetworkMap, ComplexIndsNetworkMap
from .network import TensorNetwork, add, multiply
from .packer import pack


def default_contraction_alg():
    return "bp"  # src/itensornetworkfunction.jl:16 (exact on trees; accepted and ignored here)


class Plan:
    """Owns one `ttn_plan*` (device memory, streams) for one packed network."""

    def __init__(self, packed, device=0, ngpus=1, devices=None):
        """device: the GPU of a single-device plan.  ngpus > 1 (or an explicit `devices` list): the plan is
        replicated on those GPUs of this process and every call shards its points over them
        (ttn_plan_create_multi; SURVEY 8(b),(e))."""
        self.packed = packed
        self._h = C.c_void_p()
        L = _capi.lib()
        if devices is None and int(ngpus) <= 1:
            self.device = int(device)
            self.devices = [self.device]
            _capi.check(L.ttn_plan_create(C.byref(packed.desc()), self.device, C.byref(self._h)))
        else:
            self.devices = [int(d) for d in devices] if devices is not None else list(range(int(ngpus)))
            self.device = self.devices[0]
            arr = (C.c_int32 * len(self.devices))(*self.devices)
            _capi.check(L.ttn_plan_create_multi(C.byref(packed.desc()), len(self.devices), arr, C.byref(self._h)))

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and _capi is not None and _capi._lib is not None:
            try:
                _capi._lib.ttn_plan_destroy(h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    __del__ = close

    def info(self):
        info = _capi.ttn_info()
        _capi.check(_capi.lib().ttn_plan_info(self._h, C.byref(info)))
        return {f: getattr(info, f) for f, _ in _capi.ttn_info._fields_}

    @staticmethod
    def _opts(kernel, reduce_sum, coords_mem=0, out_mem=0, chunk_points=0, accuracy=None, refine_tau=0.0,
              host_staging=0):
        o = _capi.ttn_opts()
        o.accuracy = _capi.ACCURACY_IDS[accuracy] if not isinstance(accuracy, (int, np.integer)) else int(accuracy)
        o.refine_tau = float(refine_tau)
        o.host_staging = int(host_staging)
        o.coords_mem, o.out_mem = coords_mem, out_mem
        o.kernel = _capi.KERNEL_IDS[kernel] if isinstance(kernel, str) else int(kernel or 0)
        o.reduce_sum = _capi.REDUCE_IDS[reduce_sum] if not isinstance(reduce_sum, (int, np.integer)) or isinstance(reduce_sum, bool) else int(reduce_sum)
        o.chunk_points = int(chunk_points)
        return o

    def evaluate_host(self, coords, layout=_capi.TTN_LAYOUT_AOS, kernel="auto", reduce_sum=False,
                      want_values=True, chunk_points=0, out=None, weights=None, accuracy=None, refine_tau=0.0,
                      host_staging=0):
        """coords: float64 array, (npts, n_coords) for AOS or (n_coords, npts) for SOA.  Pageable arrays
        (plain numpy) go through the library's pinned staging ring; pinned ones are used in place."""
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        nc = self.packed.n_coords
        npts = coords.shape[0] if layout == _capi.TTN_LAYOUT_AOS else coords.shape[1]
        if coords.ndim != 2 or coords.size != npts * nc:
            raise ValueError(f"coords must hold {nc} coordinate slots per point")
        dt = np.complex128 if self.packed.is_complex else np.float64
        if out is not None:  # caller-owned (e.g. pinned) result buffer
            if out.dtype != dt or out.size != npts or not out.flags.c_contiguous:
                raise ValueError("out must be a contiguous array of npts values of the network's eltype")
        elif want_values:
            out = np.empty(npts, dtype=dt)
        o = self._opts(kernel, reduce_sum, chunk_points=chunk_points, accuracy=accuracy, refine_tau=refine_tau,
                       host_staging=host_staging)
        if weights is not None:
            weights = np.ascontiguousarray(weights, dtype=np.float64)
            if weights.size != npts:
                raise ValueError("weights must hold one real weight per point")
            o.weights, o.weights_mem = weights.ctypes.data_as(C.c_void_p), _capi.TTN_MEM_HOST
        rc = _capi.lib().ttn_evaluate(
            self._h, coords.ctypes.data_as(C.c_void_p), npts, nc, layout,
            out.ctypes.data_as(C.c_void_p) if out is not None else None, C.byref(o))
        _capi.check(rc)
        return out, o

    def evaluate_device(self, coords_ptr, npts, out_ptr, layout=_capi.TTN_LAYOUT_AOS,
                        kernel="auto", reduce_sum=False, accuracy=None, refine_tau=0.0):
        """Raw device pointers (e.g. torch tensors' data_ptr()).  On a multi-GPU plan the arrays may live on any
        one GPU: the other GPUs move their blocks with peer copies."""
        o = self._opts(kernel, reduce_sum, _capi.TTN_MEM_DEVICE, _capi.TTN_MEM_DEVICE, accuracy=accuracy,
                       refine_tau=refine_tau)
        rc = _capi.lib().ttn_evaluate(self._h, C.c_void_p(coords_ptr), int(npts),
                                      self.packed.n_coords, layout,
                                      C.c_void_p(out_ptr) if out_ptr else None, C.byref(o))
        _capi.check(rc)
        return o

    def evaluate_grid(self, steps, counts, first=0, npts=None, kernel="auto", reduce_sum=True,
                      want_values=False, out_ptr=None, accuracy=None):
        steps = np.ascontiguousarray(steps, dtype=np.float64)
        counts = np.ascontiguousarray(counts, dtype=np.int64)
        total = int(np.prod(counts.astype(object)))
        if npts is None:
            npts = total - first
        g = _capi.ttn_grid()
        g.n_coords = len(steps)
        g.step = steps.ctypes.data_as(C.POINTER(C.c_double))
        g.count = counts.ctypes.data_as(C.POINTER(C.c_int64))
        g.first, g.npts = int(first), int(npts)
        out = None
        o = self._opts(kernel, reduce_sum, accuracy=accuracy)
        ptr = None
        if out_ptr is not None:
            o.out_mem = _capi.TTN_MEM_DEVICE
            ptr = C.c_void_p(out_ptr)
        elif want_values:
            out = np.empty(npts, dtype=np.complex128 if self.packed.is_complex else np.float64)
            ptr = out.ctypes.data_as(C.c_void_p)
        _capi.check(_capi.lib().ttn_evaluate_grid(self._h, C.byref(g), ptr, C.byref(o)))
        return out, o

    def evaluate_indices_host(self, index_values, kernel="auto", reduce_sum=False, want_values=True):
        """index_values: uint8 array (npts, n_sites), columns in the order of packed.site_inds."""
        iv = np.asarray(index_values)
        ns = len(self.packed.site_dim)
        if iv.ndim != 2 or iv.shape[1] != ns:
            raise ValueError(f"index_values must have {ns} columns (one per site index)")
        if iv.dtype != np.uint8:
            # range-check BEFORE the cast: 256 would wrap to a valid-looking 0
            if iv.size and (iv.min() < 0 or (iv >= np.asarray(self.packed.site_dim)[None, :]).any()):
                raise _capi.TTNError(_capi.TTN_ERR_INVALID, "an index value is out of range for its site index")
        iv = np.ascontiguousarray(iv, dtype=np.uint8)
        npts = iv.shape[0]
        out = np.empty(npts, dtype=np.complex128 if self.packed.is_complex else np.float64) if want_values else None
        o = self._opts(kernel, reduce_sum)
        _capi.check(_capi.lib().ttn_evaluate_indices(self._h, iv.ctypes.data_as(C.c_void_p), npts,
                                                     out.ctypes.data_as(C.c_void_p) if out is not None else None,
                                                     C.byref(o)))
        return out, o

    def digits_host(self, coords, layout=_capi.TTN_LAYOUT_AOS):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        nc = self.packed.n_coords
        if coords.ndim != 2:
            raise ValueError(f"coords must be a 2-D array holding {nc} coordinate slots per point")
        npts = coords.shape[0] if layout == _capi.TTN_LAYOUT_AOS else coords.shape[1]
        if coords.size != npts * nc:
            raise ValueError(f"coords must hold {nc} coordinate slots per point")
        out = np.empty((npts, len(self.packed.site_dim)), dtype=np.uint8)
        o = self._opts("auto", False)
        _capi.check(_capi.lib().ttn_digits(self._h, coords.ctypes.data_as(C.c_void_p), npts, nc,
                                           layout, out.ctypes.data_as(C.c_void_p), C.byref(o)))
        return out


class ITensorNetworkFunction:
    """src/itensornetworkfunction.jl:18-22 — {itensornetwork, indsnetworkmap}."""

    def __init__(self, itensornetwork: TensorNetwork, indsnetworkmap=None, imag_dimension_vertices=None):
        if isinstance(indsnetworkmap, IndsNetworkMap):
            inm = indsnetworkmap
        else:
            raise TypeError("indsnetworkmap must be an IndsNetworkMap")
        self.itensornetwork = itensornetwork
        self.indsnetworkmap = inm
        self._plans = {}

    @property
    def indexmap(self):
        return self.indsnetworkmap.indexmap

    def copy(self):
        return ITensorNetworkFunction(self.itensornetwork.copy(), self.indsnetworkmap.copy())

    def vertices(self):
        return self.itensornetwork.vertices()

    def __getitem__(self, v):
        return self.itensornetwork[v]

    def __setitem__(self, v, t):
        self.itensornetwork[v] = t  # bumps the network's version: cached plans of the old tensors are dropped

    def invalidate_plans(self):
        """Drop the cached GPU plans.  Needed only after mutating a tensor's ARRAY in place
        (`fitn[v].array[...] = x`); assigning a vertex (`fitn[v] = t`, `fitn.itensornetwork[v] = t`, the
        reference's `psi[v] = ...` / `psi[v] *= c`) is tracked through the network's version counter."""
        self._plans.clear()

    def is_tree(self):
        return self.itensornetwork.graph.is_tree()

    def maxlinkdim(self):
        return self.itensornetwork.maxlinkdim()

    def siteinds(self):
        return self.indsnetworkmap.indsnetwork

    # forwarded from the IndsNetworkMap (src/itensornetworkfunction.jl:60-80)
    def __getattr__(self, name):
        if name in ("ind", "dimension", "dimensions", "digit", "digits", "calculate_ind_values",
                    "calculate_p", "grid_points", "vertices_dimensions", "vertices_digits",
                    "vertex_digit", "vertex_dimension", "dimension_vertices"):
            return getattr(self.indsnetworkmap, name)
        raise AttributeError(name)

    def __add__(self, other):
        return ITensorNetworkFunction(add(self.itensornetwork, other.itensornetwork),
                                      self.indsnetworkmap)

    def __mul__(self, other):
        if isinstance(other, ITensorNetworkFunction):
            return ITensorNetworkFunction(multiply(self.itensornetwork, other.itensornetwork),
                                          self.indsnetworkmap)
        out = self.copy()  # scalar * network: scale one tensor
        v = out.vertices()[0]
        out.itensornetwork[v] = out.itensornetwork[v] * other
        return out

    __rmul__ = __mul__

    # ---- plans --------------------------------------------------------------------
    def plan(self, dims=None, device=0, ngpus=1) -> Plan:
        """The cached GPU plan of (dims, device / ngpus) for the network's CURRENT tensors.  Plans built for an
        older version of the network (any `psi[v] = t` since) are dropped first, so a mutated network is never
        evaluated with stale packed tensors; see invalidate_plans() for in-place array edits."""
        if dims is None:
            dims = self.indexmap.dimensions()
        ver = self.itensornetwork.version
        if getattr(self, "_plans_version", ver) != ver:
            self._plans.clear()
        self._plans_version = ver
        key = (tuple(int(d) for d in dims), int(device), int(ngpus))
        if key not in self._plans:
            self._plans[key] = Plan(pack(self, list(key[0])), device=device, ngpus=ngpus)
        return self._plans[key]


def _points_to_coords(fitn, xs, dims):
    """Normalise the reference's argument forms.  Returns (coords[npts, n_coords], dims, single)."""
    imap = fitn.indexmap
    is_cmap = isinstance(imap, ComplexIndexMap)
    single = False
    if np.isscalar(xs):
        xs = [[xs]]
        single = True
        if dims is None:
            dims = [imap.dimensions()[0]]  # first(dimensions(fitn)), :108-112
        elif np.isscalar(dims):
            dims = [dims]
    else:
        arr = np.asarray(xs)
        if arr.ndim == 1:
            xs = [list(xs)]
            single = True
    pts = np.asarray(xs)
    if pts.ndim != 2:
        raise ValueError("points must be a (npts, D) array or a list of equal-length points")
    if dims is None:
        dims = imap.dimensions()  # default dims = dimensions(fitn), :99
    dims = [int(d) for d in dims]
    if pts.shape[1] != len(dims):
        raise AssertionError("length(xs) == length(dims)")  # realindexmap.jl:68
    if is_cmap:
        z = pts.astype(np.complex128)
        coords = np.empty((z.shape[0], 2 * z.shape[1]))
        coords[:, 0::2] = z.real
        coords[:, 1::2] = z.imag
    else:
        if np.iscomplexobj(pts):
            raise TypeError("complex coordinates need a ComplexIndexMap")
        coords = pts.astype(np.float64)
    return coords, dims, single


def evaluate(fitn: ITensorNetworkFunction, xs, dims=None, *, alg=None, device=0, ngpus=1, kernel="auto",
             reduce=None, weights=None, accuracy=None, return_opts=False):
    """Evaluate `fitn` at one point or at a batch of points (see module docstring).

    reduce=None  -> values;  reduce="sum" -> the sum over all points (grid quadrature,
    cf. integrate(...; take_sum=true), src/integration.jl:6-17); reduce="abs2" -> sum |f|^2;
    reduce="weighted" -> sum_p weights[p] * f(p) (fused quadrature functionals).
    `alg` is accepted for signature compatibility ("bp" and "exact" coincide on trees).
    ngpus=G shards the points over G GPUs of this process (contiguous blocks, network replicated; SURVEY 8 e).
    accuracy="refined" re-evaluates the points that cancel (|f| << rms) in double-double (TTN_ACCURACY_REFINED).
    """
    coords, dims, single = _points_to_coords(fitn, xs, dims)
    plan = fitn.plan(dims, device=device, ngpus=ngpus)
    out, o = plan.evaluate_host(coords, kernel=kernel, reduce_sum=reduce, want_values=(reduce is None),
                                weights=weights, accuracy=accuracy)
    if reduce == "abs2":
        res = o.sum_out[0]
    elif reduce is not None:
        res = complex(o.sum_out[0], o.sum_out[1]) if plan.packed.is_complex else o.sum_out[0]
    elif single:
        res = out[0].item()
    else:
        res = out
    return (res, o) if return_opts else res


def evaluate_indices(fitn: ITensorNetworkFunction, ind_to_ind_value_maps, *, device=0, kernel="auto"):
    """Batched project() + scalar() at given index settings (the inner loop of TCI).
    `ind_to_ind_value_maps`: a list of {Index: value} dictionaries (what calculate_ind_values
    returns), or a uint8 array (npts, n_sites) whose columns follow plan.packed.site_inds."""
    plan = fitn.plan(device=device)
    sites = plan.packed.site_inds
    if isinstance(ind_to_ind_value_maps, np.ndarray):
        iv = ind_to_ind_value_maps
    else:
        iv = np.array([[m[i] for i in sites] for m in ind_to_ind_value_maps], dtype=np.uint8).reshape(-1, len(sites))
    return plan.evaluate_indices_host(iv, kernel=kernel)[0]


def batched_ind_values(fitn, xs, dims=None, device=0):
    """Batched calculate_ind_values on the GPU: returns (digits[npts, n_sites], site_inds)."""
    coords, dims, _ = _points_to_coords(fitn, xs, dims)
    plan = fitn.plan(dims, device=device)
    return plan.digits_host(coords), plan.packed.site_inds


__all__ = ["ITensorNetworkFunction", "evaluate", "evaluate_indices", "batched_ind_values", "Plan",
           "default_contraction_alg", "RealIndsNetworkMap", "ComplexIndsNetworkMap"]
