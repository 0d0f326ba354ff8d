"""ctypes binding of libsemcode_ivf.so (the C ABI in include/semcode_ivf.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
Pointers may come from torch tensors (CPU or CUDA), numpy arrays, or be None.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

try:  # torch is plumbing (device memory + streams); the library itself does not need it
    import torch
except Exception:  # pragma: no cover
    torch = None

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsemcode_ivf.so")

SC_OK = 0
METRIC_IP = 0
METRIC_L2 = 1
MAX_K = 2048

# every symbol include/semcode_ivf.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "sc_last_error", "sc_abi_version", "sc_device_count", "sc_index_create", "sc_index_destroy", "sc_index_reset",
    "sc_index_train", "sc_index_kmeans_init", "sc_index_kmeans_step", "sc_index_kmeans_update",
    "sc_index_set_centroids", "sc_index_get_centroids", "sc_index_assign", "sc_index_probe", "sc_index_add",
    "sc_index_add_preassigned", "sc_index_remove_ids", "sc_index_search", "sc_index_search_preassigned",
    "sc_merge_topk", "sc_index_stats", "sc_index_list_sizes", "sc_index_export_list", "sc_index_export_lists",
    "sc_index_compact", "sc_index_set_profiling",
    "sc_index_last_search_times", "sc_index_set_param",
    "sc_exchange_create", "sc_exchange_destroy", "sc_exchange_status", "sc_exchange_poll", "sc_exchange_set_timeout_ms",
    "sc_index_search_sharded",
]


class NativeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsemcode_ivf error {code}: {msg}")
        self.code = code


class ScFilter(C.Structure):
    _fields_ = [
        ("repo_tags", C.POINTER(C.c_uint32)),
        ("n_repos", C.c_int32),
        ("lang_tags", C.POINTER(C.c_uint8)),
        ("n_langs", C.c_int32),
    ]


class ScStats(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("dim_padded", C.c_int32), ("metric", C.c_int32), ("nlist", C.c_int32),
        ("device", C.c_int32), ("trained", C.c_int32),
        ("ntotal", C.c_int64), ("nremoved", C.c_int64), ("npages", C.c_int64), ("nfree_pages", C.c_int64),
        ("bytes_lists", C.c_int64),
        ("bytes_scratch", C.c_int64),
        ("max_list_len", C.c_int32), ("min_list_len", C.c_int32),
    ]


class ScSearchTimes(C.Structure):
    _fields_ = [
        ("coarse_ms", C.c_float), ("probe_select_ms", C.c_float), ("plan_ms", C.c_float), ("scan_ms", C.c_float),
        ("topk_ms", C.c_float), ("total_ms", C.c_float),
        ("scanned_rows", C.c_int64), ("unique_rows", C.c_int64),
        ("scan_launches", C.c_int32), ("total_launches", C.c_int32),
    ]


_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(semcode_b200 has no CPU or PyTorch fallback)"
        )
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.sc_last_error.restype = C.c_char_p
    L.sc_last_error.argtypes = []
    L.sc_abi_version.restype = C.c_int
    L.sc_device_count.restype = C.c_int
    sig = {
        "sc_index_create": [i32, i32, i32, i32, C.POINTER(vp)],
        "sc_index_destroy": [vp],
        "sc_index_reset": [vp],
        "sc_index_train": [vp, vp, i64, i32, vp, vp, vp],
        "sc_index_kmeans_init": [vp, vp, i64, vp, vp],
        "sc_index_kmeans_step": [vp, vp, i64, vp, vp, vp, vp],
        "sc_index_kmeans_update": [vp, vp, vp, vp, vp],
        "sc_index_set_centroids": [vp, vp, i32, vp],
        "sc_index_get_centroids": [vp, vp, vp],
        "sc_index_assign": [vp, vp, i64, vp, vp],
        "sc_index_probe": [vp, vp, i64, i32, vp, vp, vp],
        "sc_index_add": [vp, vp, vp, vp, vp, i64, vp],
        "sc_index_add_preassigned": [vp, vp, vp, vp, vp, vp, i64, vp],
        "sc_index_remove_ids": [vp, vp, i64, vp, vp],
        "sc_index_search": [vp, vp, i64, i32, i32, C.POINTER(ScFilter), vp, vp, vp],
        "sc_index_search_preassigned": [vp, vp, i64, i32, i32, vp, C.POINTER(ScFilter), vp, vp, vp],
        "sc_merge_topk": [vp, vp, i32, i64, i32, i32, i32, vp, vp, i32, vp],
        "sc_index_stats": [vp, C.POINTER(ScStats)],
        "sc_index_list_sizes": [vp, vp],
        "sc_index_export_list": [vp, i32, i64, vp, vp, vp, vp, vp],
        "sc_index_export_lists": [vp, i32, i32, i64, vp, vp, vp, vp, vp],
        "sc_index_compact": [vp, vp, vp],
        "sc_index_set_profiling": [vp, i32],
        "sc_index_last_search_times": [vp, C.POINTER(ScSearchTimes)],
        "sc_index_set_param": [vp, C.c_char_p, i64],
        "sc_exchange_create": [i32, i32, C.POINTER(vp), i64, i32, C.POINTER(vp)],
        "sc_exchange_destroy": [vp],
        "sc_exchange_status": [vp, C.POINTER(i32), C.POINTER(i64)],
        "sc_exchange_poll": [vp, C.POINTER(i32)],
        "sc_exchange_set_timeout_ms": [vp, i64],
        "sc_index_search_sharded": [vp, vp, vp, i64, i32, i32, vp, C.POINTER(ScFilter), vp, vp, vp],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != SC_OK:
        msg = lib().sc_last_error()
        raise NativeError(rc, msg.decode("utf-8", "replace") if msg else "unknown error")


_NP = {
    "f32": np.float32, "f64": np.float64, "i32": np.int32, "i64": np.int64, "u32": np.uint32, "u8": np.uint8,
}


def _torch_dtype(kind: str):
    return {
        "f32": torch.float32, "f64": torch.float64, "i32": torch.int32, "i64": torch.int64, "u8": torch.uint8,
        # torch has no first-class uint32 arithmetic; int32 storage carries the same bits
        "u32": torch.int32,
    }[kind]


def ptr(obj, kind: str, *, name: str = "buffer") -> Optional[int]:
    """Raw address of a contiguous torch tensor / numpy array of element type `kind`, or None."""
    if obj is None:
        return None
    if torch is not None and isinstance(obj, torch.Tensor):
        want = _torch_dtype(kind)
        ok = obj.dtype == want or (kind == "u32" and obj.dtype == getattr(torch, "uint32", None))
        if not ok:
            raise TypeError(f"{name}: expected torch dtype {want}, got {obj.dtype}")
        if not obj.is_contiguous():
            raise ValueError(f"{name}: tensor must be contiguous")
        return obj.data_ptr() if obj.numel() else None
    if isinstance(obj, np.ndarray):
        if obj.dtype != _NP[kind]:
            raise TypeError(f"{name}: expected numpy dtype {_NP[kind].__name__}, got {obj.dtype}")
        if not obj.flags.c_contiguous:
            raise ValueError(f"{name}: array must be C-contiguous")
        return obj.ctypes.data if obj.size else None
    raise TypeError(f"{name}: expected a torch tensor or numpy array, got {type(obj).__name__}")


def current_stream(device: int) -> int:
    """cudaStream_t of torch's current stream on `device` (0 = legacy default stream)."""
    if torch is None or not torch.cuda.is_available():
        return 0
    return int(torch.cuda.current_stream(device).cuda_stream)
