"""ctypes binding of oracle/liboracle.so (the CPU restatement).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from tgsfilter_b200 import _capi
from tgsfilter_b200.params import FilterParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(ORACLE_DIR, "tgsf_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    lib = C.CDLL(LIB)
    lib.tgsfo_align_hw.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                   C.POINTER(_capi.AlignResult), C.c_void_p, C.c_void_p, C.c_int]
    lib.tgsfo_counter_layout.argtypes = [C.POINTER(_capi.Params), C.POINTER(_capi.CounterLayout)]
    lib.tgsfo_run.argtypes = [C.POINTER(_capi.Params), C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                              C.POINTER(C.c_uint32), C.c_void_p]
    lib.tgsfo_base_content_counts.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.tgsfo_base_content_counts.restype = None
    lib.tgsfo_base_content_trim.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
    lib.tgsfo_adapter_search.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32,
                                         C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.c_int32,
                                         C.c_float, C.c_void_p]
    lib.tgsfo_adapter_search.restype = None
    _lib = lib
    return lib


def align_hw(q: bytes, t: bytes, k: int, loc_cap: int = 4096):
    """Returns (AlignResult-as-dict, [(start, end), ...])."""
    lib = load()
    res = _capi.AlignResult()
    starts = np.zeros(loc_cap, dtype=np.int32)
    ends = np.zeros(loc_cap, dtype=np.int32)
    qb = np.frombuffer(q, dtype=np.uint8) if len(q) else np.zeros(1, np.uint8)
    tb = np.frombuffer(t, dtype=np.uint8) if len(t) else np.zeros(1, np.uint8)
    lib.tgsfo_align_hw(qb.ctypes.data, len(q), tb.ctypes.data, len(t), k, C.byref(res),
                       starts.ctypes.data, ends.ctypes.data, loc_cap)
    n = min(res.n_locations, loc_cap)
    d = {f: getattr(res, f) for f, _ in _capi.AlignResult._fields_}
    return d, list(zip(starts[:n].tolist(), ends[:n].tolist()))


def layout(params: FilterParams) -> _capi.CounterLayout:
    lib = load()
    p, keep = params.to_c()
    L = _capi.CounterLayout()
    lib.tgsfo_counter_layout(C.byref(p), C.byref(L))
    return L


def run(params: FilterParams, batch, counters: np.ndarray | None = None):
    """Oracle equivalent of submit+collect+counters.  Returns (reads, pieces, counters)."""
    lib = load()
    p, keep = params.to_c()
    L = layout(params)
    if counters is None:
        counters = np.zeros(L.n_u64, dtype=np.uint64)
    n = batch.n_reads
    reads = np.zeros(n, dtype=_capi.READ_RESULT_DTYPE)
    cap = max(16, 4 * n)
    while True:
        pieces = np.zeros(cap, dtype=_capi.PIECE_DTYPE)
        npieces = C.c_uint32(0)
        cnt = counters.copy()
        bases = np.ascontiguousarray(batch.bases)
        quals = None if batch.quals is None else np.ascontiguousarray(batch.quals)
        offs = np.ascontiguousarray(batch.offsets, dtype=np.uint64)
        rc = lib.tgsfo_run(C.byref(p), bases.ctypes.data,
                           None if quals is None else quals.ctypes.data, offs.ctypes.data, n,
                           reads.ctypes.data, pieces.ctypes.data, cap, C.byref(npieces),
                           cnt.ctypes.data)
        if rc == _capi.TGSF_ERR_CAPACITY and npieces.value > cap:
            cap = npieces.value
            continue
        assert rc == 0, rc
        return reads, pieces[:npieces.value], cnt


def base_content_counts(ends: np.ndarray) -> np.ndarray:
    lib = load()
    n, row = ends.shape
    out = np.zeros((row, 4), dtype=np.int32)
    ends = np.ascontiguousarray(ends)
    lib.tgsfo_base_content_counts(ends.ctypes.data, n, row, out.ctypes.data)
    return out


def base_content_trim(bases_num: np.ndarray, seq_num: int, end_bias: float) -> int:
    lib = load()
    bn = np.ascontiguousarray(bases_num, dtype=np.int32)
    return lib.tgsfo_base_content_trim(bn.ctypes.data, bn.shape[0], seq_num, end_bias)


def adapter_search(ends: np.ndarray, lib_seqs, min_sim: float) -> np.ndarray:
    lib = load()
    n, row = ends.shape
    ends = np.ascontiguousarray(ends)
    seqs = (C.c_char_p * len(lib_seqs))(*lib_seqs)
    lens = (C.c_int32 * len(lib_seqs))(*[len(s) for s in lib_seqs])
    maps = np.zeros(len(lib_seqs), dtype=np.int64)
    lib.tgsfo_adapter_search(ends.ctypes.data, n, row, seqs, lens, len(lib_seqs), min_sim,
                             maps.ctypes.data)
    return maps
