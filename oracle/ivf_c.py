"""ctypes binding of oracle/liborc.so (C restatement; TEST INFRASTRUCTURE ONLY, parity unpinned).

Used by tests/ as a second independent checker and by bench.py as the timed CPU baseline
(cpu_baseline.kind == "port": the reference's engine cannot be built here).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")
_lib = None


def _cpu_key() -> str:
    """Identifies the host CPU (model + ISA flags): a -march=native build is only valid on the CPU it was made on."""
    import hashlib

    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        model = next((l for l in txt.splitlines() if l.startswith("model name")), "")
        flags = next((l for l in txt.splitlines() if l.startswith("flags")), "")
    except OSError:
        model, flags = "unknown", ""
    return hashlib.sha1((model + flags).encode()).hexdigest()[:12]


def build(force: bool = False, native: bool = False) -> str:
    """Compile oracle/liborc.so (portable x86-64-v3).  native=True (bench.py's CPU arm): compile for THIS host with
    -march=native into oracle/_native/ (git- and gpurun-ignored, keyed by the CPU, never shipped) and bind that instead;
    falls back to the portable build if the compile fails."""
    global _LIB_PATH, _lib
    src = os.path.join(_HERE, "ivf_oracle.c")
    if native:
        out = os.path.join(_HERE, "_native", f"liborc-{_cpu_key()}.so")
        try:
            if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
                os.makedirs(os.path.dirname(out), exist_ok=True)
                cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
                subprocess.run([cc, "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
                                "-o", out + ".tmp", src, "-lm"], check=True, capture_output=True)
                os.replace(out + ".tmp", out)
            if _LIB_PATH != out:
                _LIB_PATH, _lib = out, None
            return out
        except (subprocess.CalledProcessError, OSError):
            pass
    portable = os.path.join(_HERE, "liborc.so")
    if force or not os.path.exists(portable) or os.path.getmtime(portable) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liborc.so"], check=True, capture_output=True)
    return _LIB_PATH


def is_native() -> bool:
    return os.sep + "_native" + os.sep in _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        f32p, i32p, i64p, u8p = (C.POINTER(t) for t in (C.c_float, C.c_int32, C.c_int64, C.c_uint8))
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_set_num_threads.restype = None
        L.orc_gemm_nt.argtypes = [f32p, C.c_int64, f32p, C.c_int, C.c_int, f32p]
        L.orc_coarse_scores.argtypes = [f32p, C.c_int64, f32p, C.c_int, C.c_int, C.c_int, f32p]
        L.orc_top_probes.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, i32p]
        L.orc_assign.argtypes = [f32p, C.c_int64, f32p, C.c_int, C.c_int, C.c_int, i32p]
        L.orc_scan_search.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, i32p, C.c_int, i64p, f32p, i64p, u8p,
                                      C.c_int, f32p, i64p]
        L.orc_search.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, f32p, C.c_int, C.c_int, i64p, f32p, i64p, u8p,
                                 C.c_int, f32p, i64p]
        for fn in (L.orc_gemm_nt, L.orc_coarse_scores, L.orc_top_probes, L.orc_assign, L.orc_scan_search,
                   L.orc_search):
            fn.restype = None
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray], ctype):
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    return a.ctypes.data_as(C.POINTER(ctype))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def use_all_cores() -> int:
    """OpenMP threads = the cores this process may run on (a launcher such as torchrun sets OMP_NUM_THREADS=1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(n)
    return num_threads()


def coarse_scores(q, centroids, metric: int) -> np.ndarray:
    q, c = _f32(q), _f32(centroids)
    out = np.empty((q.shape[0], c.shape[0]), dtype=np.float32)
    lib().orc_coarse_scores(_p(q, C.c_float), q.shape[0], _p(c, C.c_float), c.shape[0], c.shape[1], int(metric),
                            _p(out, C.c_float))
    return out


def top_probes(scores: np.ndarray, nprobe: int) -> np.ndarray:
    scores = _f32(scores)
    nq, nlist = scores.shape
    nprobe = min(nprobe, nlist)
    out = np.empty((nq, nprobe), dtype=np.int32)
    lib().orc_top_probes(_p(scores, C.c_float), nq, nlist, nprobe, _p(out, C.c_int32))
    return out


def assign(x, centroids, metric: int) -> np.ndarray:
    x, c = _f32(x), _f32(centroids)
    out = np.empty(x.shape[0], dtype=np.int32)
    lib().orc_assign(_p(x, C.c_float), x.shape[0], _p(c, C.c_float), c.shape[0], c.shape[1], int(metric),
                     _p(out, C.c_int32))
    return out


def scan_search(q, metric: int, probes, list_off, vecs, ids, k: int, skip=None) -> Tuple[np.ndarray, np.ndarray]:
    q, vecs = _f32(q), _f32(vecs)
    probes = np.ascontiguousarray(probes, dtype=np.int32)
    list_off = np.ascontiguousarray(list_off, dtype=np.int64)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    if skip is not None:
        skip = np.ascontiguousarray(skip, dtype=np.uint8)
    nq, d = q.shape
    out_d = np.empty((nq, k), dtype=np.float32)
    out_i = np.empty((nq, k), dtype=np.int64)
    lib().orc_scan_search(_p(q, C.c_float), nq, d, int(metric), _p(probes, C.c_int32), probes.shape[1],
                          _p(list_off, C.c_int64), _p(vecs, C.c_float), _p(ids, C.c_int64), _p(skip, C.c_uint8), k,
                          _p(out_d, C.c_float), _p(out_i, C.c_int64))
    return out_d, out_i


def search(q, metric: int, centroids, nprobe: int, list_off, vecs, ids, k: int, skip=None):
    """coarse + probes + scan in one timed call (CPU baseline)."""
    q, vecs, c = _f32(q), _f32(vecs), _f32(centroids)
    list_off = np.ascontiguousarray(list_off, dtype=np.int64)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    if skip is not None:
        skip = np.ascontiguousarray(skip, dtype=np.uint8)
    nq, d = q.shape
    out_d = np.empty((nq, k), dtype=np.float32)
    out_i = np.empty((nq, k), dtype=np.int64)
    lib().orc_search(_p(q, C.c_float), nq, d, int(metric), _p(c, C.c_float), c.shape[0], nprobe,
                     _p(list_off, C.c_int64), _p(vecs, C.c_float), _p(ids, C.c_int64), _p(skip, C.c_uint8), k,
                     _p(out_d, C.c_float), _p(out_i, C.c_int64))
    return out_d, out_i
