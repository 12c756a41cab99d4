"""ctypes binding of include/ttneval.h (libttneval.so).  There is no fallback: if the CUDA
library has not been built, importing the binding raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LIBTTNEVAL") or os.path.join(_HERE, "csrc", "libttneval.so")   # same override as the Julia wrapper

TTN_ABI_VERSION = 3
TTN_OK, TTN_ERR_INVALID, TTN_ERR_DOMAIN, TTN_ERR_CUDA, TTN_ERR_UNSUPPORTED, TTN_ERR_NOMEM = range(6)
TTN_LAYOUT_AOS, TTN_LAYOUT_SOA = 0, 1
TTN_MEM_HOST, TTN_MEM_DEVICE = 0, 1
TTN_KERNEL_AUTO, TTN_KERNEL_GENERIC, TTN_KERNEL_CHAIN, TTN_KERNEL_DMMA, TTN_KERNEL_GEMM, TTN_KERNEL_TREE, TTN_KERNEL_GRID, TTN_KERNEL_TABLE = range(8)
KERNEL_NAMES = {0: "auto", 1: "generic", 2: "chain", 3: "dmma", 4: "gemm", 5: "tree", 6: "grid", 7: "table"}
KERNEL_IDS = {v: k for k, v in KERNEL_NAMES.items()}
TTN_REDUCE_NONE, TTN_REDUCE_SUM, TTN_REDUCE_ABS2, TTN_REDUCE_WEIGHTED = range(4)
REDUCE_IDS = {None: 0, False: 0, "none": 0, True: 1, "sum": 1, "abs2": 2, "weighted": 3}
TTN_STAGE_AUTO, TTN_STAGE_OFF, TTN_STAGE_COPY = 0, 1, 2
TTN_ACCURACY_FP64, TTN_ACCURACY_REFINED = 0, 1
ACCURACY_IDS = {None: 0, "fp64": 0, "refined": 1}


class ttn_desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("n_vertices", C.c_int32),
        ("n_coords", C.c_int32),
        ("is_complex", C.c_int32),
        ("root", C.c_int32),
        ("n_sites", C.c_int32),
        ("parent", C.POINTER(C.c_int32)),
        ("link_dim", C.POINTER(C.c_int32)),
        ("site_ptr", C.POINTER(C.c_int32)),
        ("site_dim", C.POINTER(C.c_int32)),
        ("site_coord", C.POINTER(C.c_int32)),
        ("site_digit", C.POINTER(C.c_int32)),
        ("thr_ptr", C.POINTER(C.c_int32)),
        ("thr", C.POINTER(C.c_double)),
        ("tensor_ptr", C.POINTER(C.c_int64)),
        ("tensors", C.c_void_p),
    ]


class ttn_opts(C.Structure):
    _fields_ = [
        ("coords_mem", C.c_int32),
        ("out_mem", C.c_int32),
        ("kernel", C.c_int32),
        ("reduce_sum", C.c_int32),
        ("chunk_points", C.c_int64),
        ("sum_out", C.c_double * 2),
        ("kernel_ms", C.c_float),
        ("total_ms", C.c_float),
        ("kernel_used", C.c_int32),
        ("n_launches", C.c_int32),
        ("weights", C.c_void_p),
        ("weights_mem", C.c_int32),
        ("reserved_", C.c_int32),
        ("flops_executed", C.c_double),
        ("host_staging", C.c_int32),
        ("accuracy", C.c_int32),
        ("refine_tau", C.c_double),
        ("n_devices_used", C.c_int32),
        ("staged", C.c_int32),
        ("n_refined", C.c_int64),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
    ]


class ttn_grid(C.Structure):
    _fields_ = [
        ("n_coords", C.c_int32),
        ("step", C.POINTER(C.c_double)),
        ("count", C.POINTER(C.c_int64)),
        ("first", C.c_int64),
        ("npts", C.c_int64),
    ]


class ttn_info(C.Structure):
    _fields_ = [
        ("n_vertices", C.c_int32),
        ("n_coords", C.c_int32),
        ("is_complex", C.c_int32),
        ("n_sites", C.c_int32),
        ("max_link_dim", C.c_int32),
        ("is_chain", C.c_int32),
        ("auto_kernel", C.c_int32),
        ("device", C.c_int32),
        ("kernels_available", C.c_int32),
        ("n_devices", C.c_int32),
        ("flops_per_point", C.c_double),
        ("bytes_per_point", C.c_double),
        ("tensor_bytes", C.c_int64),
    ]


EXPORTS = [
    "ttn_plan_create", "ttn_plan_create_multi", "ttn_host_register", "ttn_host_unregister", "ttn_plan_destroy", "ttn_plan_info", "ttn_evaluate", "ttn_evaluate_grid",
    "ttn_evaluate_indices",
    "ttn_digits", "ttn_measure_fp64_peak", "ttn_last_error", "ttn_device_count",
    "ttn_abi_version",
]

_lib = None


class TTNError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libttneval error {code}: {msg}")
        self.code = code


def lib():
    """Load libttneval.so (once).  Raises if it is missing — there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  The batched evaluate path has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.ttn_plan_create.argtypes = [C.POINTER(ttn_desc), C.c_int32, C.POINTER(vp)]
    L.ttn_plan_create.restype = C.c_int
    L.ttn_plan_create_multi.argtypes = [C.POINTER(ttn_desc), C.c_int32, C.POINTER(C.c_int32), C.POINTER(vp)]
    L.ttn_plan_create_multi.restype = C.c_int
    L.ttn_host_register.argtypes = [vp, C.c_uint64]
    L.ttn_host_register.restype = C.c_int
    L.ttn_host_unregister.argtypes = [vp]
    L.ttn_host_unregister.restype = C.c_int
    L.ttn_plan_destroy.argtypes = [vp]
    L.ttn_plan_destroy.restype = None
    L.ttn_plan_info.argtypes = [vp, C.POINTER(ttn_info)]
    L.ttn_plan_info.restype = C.c_int
    L.ttn_evaluate.argtypes = [vp, vp, C.c_int64, C.c_int32, C.c_int32, vp, C.POINTER(ttn_opts)]
    L.ttn_evaluate.restype = C.c_int
    L.ttn_evaluate_grid.argtypes = [vp, C.POINTER(ttn_grid), vp, C.POINTER(ttn_opts)]
    L.ttn_evaluate_grid.restype = C.c_int
    L.ttn_evaluate_indices.argtypes = [vp, vp, C.c_int64, vp, C.POINTER(ttn_opts)]
    L.ttn_evaluate_indices.restype = C.c_int
    L.ttn_digits.argtypes = [vp, vp, C.c_int64, C.c_int32, C.c_int32, vp, C.POINTER(ttn_opts)]
    L.ttn_digits.restype = C.c_int
    L.ttn_measure_fp64_peak.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ttn_measure_fp64_peak.restype = C.c_int
    L.ttn_last_error.argtypes = []
    L.ttn_last_error.restype = C.c_char_p
    L.ttn_device_count.argtypes = []
    L.ttn_device_count.restype = C.c_int
    L.ttn_abi_version.argtypes = []
    L.ttn_abi_version.restype = C.c_int
    if L.ttn_abi_version() != TTN_ABI_VERSION:
        raise RuntimeError("libttneval.so ABI version mismatch; rebuild")
    _lib = L
    return L


def check(rc):
    if rc != TTN_OK:
        msg = lib().ttn_last_error()
        raise TTNError(rc, msg.decode() if msg else "")
