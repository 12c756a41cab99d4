"""Python face of the CPU oracle (TEST INFRASTRUCTURE ONLY — see ttn_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (itensornumericalanalysis.jl_b200) never does.

 - `digits`, `evaluate`   ctypes calls into libttn_oracle.so (C restatement, takes the same
                           flat `ttn_desc` the CUDA library takes)
 - `dense_evaluate`        an independent numpy check that does NOT go through the packer: the
                           whole network is contracted to the dense b^L tensor (the analogue of
                           build_full_rank_tensor, src/utils.jl:28-39) and indexed with the
                           digits from the host-side calculate_ind_values.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import itna_b200  # noqa: E402  (struct definitions + host mirror only)
from itna_b200 import _capi  # noqa: E402

ORACLE_LD, ORACLE_F64, ORACLE_BP = 0, 1, 2
_LIB = os.path.join(_HERE, "libttn_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
            for f in ("ttn_oracle.c", "ttn_oracle_body.inc")):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libttn_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        vp = C.c_void_p
        L.oracle_digits.argtypes = [C.POINTER(_capi.ttn_desc), vp, C.c_int64, C.c_int32, vp]
        L.oracle_digits.restype = C.c_int
        L.oracle_evaluate.argtypes = [C.POINTER(_capi.ttn_desc), vp, C.c_int64, C.c_int32,
                                      C.c_int32, C.c_int32, vp]
        L.oracle_evaluate.restype = C.c_int
        L.oracle_evaluate_ld2.argtypes = [C.POINTER(_capi.ttn_desc), vp, C.c_int64, C.c_int32, vp, vp]
        L.oracle_evaluate_ld2.restype = C.c_int
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle error {code}")
        self.code = code


def _coords(packed, coords, layout):
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    npts = coords.shape[0] if layout == _capi.TTN_LAYOUT_AOS else coords.shape[1]
    assert coords.size == npts * packed.n_coords
    return coords, npts


def digits(packed, coords, layout=_capi.TTN_LAYOUT_AOS):
    coords, npts = _coords(packed, coords, layout)
    out = np.empty((npts, len(packed.site_dim)), dtype=np.uint8)
    rc = lib().oracle_digits(C.byref(packed.desc()), coords.ctypes.data_as(C.c_void_p), npts,
                             layout, out.ctypes.data_as(C.c_void_p))
    if rc:
        raise OracleError(rc)
    return out


def evaluate(packed, coords, mode=ORACLE_LD, nthreads=1, layout=_capi.TTN_LAYOUT_AOS):
    coords, npts = _coords(packed, coords, layout)
    out = np.empty(npts, dtype=np.complex128 if packed.is_complex else np.float64)
    rc = lib().oracle_evaluate(C.byref(packed.desc()), coords.ctypes.data_as(C.c_void_p), npts,
                               layout, mode, nthreads, out.ctypes.data_as(C.c_void_p))
    if rc:
        raise OracleError(rc)
    return out


def evaluate_ld2(packed, coords, layout=_capi.TTN_LAYOUT_AOS):
    """80-bit result as (hi, lo) double pairs: hi + lo carries the full 64-bit mantissa."""
    coords, npts = _coords(packed, coords, layout)
    dt = np.complex128 if packed.is_complex else np.float64
    hi, lo = np.empty(npts, dtype=dt), np.empty(npts, dtype=dt)
    rc = lib().oracle_evaluate_ld2(C.byref(packed.desc()), coords.ctypes.data_as(C.c_void_p),
                                   npts, layout, hi.ctypes.data_as(C.c_void_p),
                                   lo.ctypes.data_as(C.c_void_p))
    if rc:
        raise OracleError(rc)
    return hi, lo


def max_threads():
    return int(lib().oracle_max_threads())


# ----------------------------------------------------------------------------- numpy check


def dense_tensor(fitn):
    """Contract the whole network over its link indices -> (dense array, site inds)."""
    tn = fitn.itensornetwork
    ids = {}

    def sym(ind):
        if ind not in ids:
            ids[ind] = len(ids)
        return ids[ind]

    operands = []
    site_inds = []
    for v in tn.vertices():
        t = tn[v]
        operands += [t.array, [sym(i) for i in t.inds]]
        site_inds += [i for i in fitn.indsnetworkmap[v]]
    out = [sym(i) for i in site_inds]
    return np.einsum(*operands, out, optimize=True), site_inds


def dense_evaluate(fitn, points, dims=None):
    """Evaluate by indexing the dense tensor with host-side calculate_ind_values digits."""
    dense, site_inds = dense_tensor(fitn)
    vals = []
    for p in points:
        p = list(p) if not np.isscalar(p) else p
        m = fitn.indsnetworkmap.calculate_ind_values(p, dims)
        vals.append(dense[tuple(m[i] for i in site_inds)])
    return np.asarray(vals)


def error_metric(v, ref):
    """|v - ref| / max(|ref|, 1e-3 * rms(ref)), and the unfloored relative error."""
    v, ref = np.asarray(v), np.asarray(ref)
    rms = np.sqrt(np.mean(np.abs(ref) ** 2)) if ref.size else 0.0
    floor = np.maximum(np.abs(ref), 1e-3 * rms)
    floor = np.where(floor == 0, 1.0, floor)
    return np.abs(v - ref) / floor
