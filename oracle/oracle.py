"""ctypes handle on oracle/libmash_oracle.so (CPU restatement of Mash).

TEST INFRASTRUCTURE ONLY -- see the header of mash_oracle.c.  Importers:
tests/, bench.py (cpu_baseline / --impl reference) and __graft_entry__.smoke().
The product package `auriclass_b200` must never import this module.
"""
from __future__ import annotations

import ctypes as C
import gzip
import subprocess
from dataclasses import dataclass
from pathlib import Path
from typing import List, Sequence, Tuple

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libmash_oracle.so"


class _Stats(C.Structure):
    _fields_ = [
        ("n_records", C.c_uint64), ("n_kept", C.c_uint64), ("total_len", C.c_uint64),
        ("n_kmers", C.c_uint64), ("n_distinct", C.c_uint64), ("n_qualified", C.c_uint64),
        ("n_sketch", C.c_uint64), ("genome_size", C.c_double), ("coverage", C.c_double),
        ("error", C.c_int),
    ]


class _Dist(C.Structure):
    _fields_ = [("common", C.c_uint64), ("denom", C.c_uint64),
                ("distance", C.c_double), ("pvalue", C.c_double)]


def build(force: bool = False) -> Path:
    src = _HERE / "mash_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "libmash_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        u8pp, szp = C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)
        for name in ("orc_sketch", "orc_sketch_stream"):
            f = getattr(L, name)
            f.restype = C.c_int64
            f.argtypes = [u8pp, szp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
                          C.c_void_p, C.c_void_p, C.POINTER(_Stats)]
        L.orc_sketch_parallel.restype = C.c_int64
        L.orc_sketch_parallel.argtypes = [u8pp, szp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                          C.c_int, C.c_void_p, C.c_void_p, C.POINTER(_Stats)]
        L.orc_all_hashes.restype = C.c_int64
        L.orc_all_hashes.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_void_p, C.c_size_t]
        L.orc_murmur3_x64_128.restype = None
        L.orc_murmur3_x64_128.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.POINTER(C.c_uint64 * 2)]
        L.orc_dist.restype = None
        L.orc_dist.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int,
                               C.c_uint64, C.c_uint64, C.POINTER(_Dist)]
        L.orc_binomial_Q.restype = C.c_double
        L.orc_binomial_Q.argtypes = [C.c_uint64, C.c_double, C.c_uint64]
        L.orc_bounds_text.restype = C.c_int
        L.orc_bounds_text.argtypes = [C.c_int, C.c_double, C.c_char_p, C.c_size_t]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def read_maybe_gz(path) -> bytes:
    raw = Path(path).read_bytes()
    return gzip.decompress(raw) if raw[:2] == b"\x1f\x8b" else raw


@dataclass
class OracleSketch:
    hashes: np.ndarray
    counts: np.ndarray
    n_records: int
    n_kept: int
    total_len: int
    n_kmers: int
    n_distinct: int
    n_qualified: int
    genome_size: float
    coverage: float


def _call_sketch(fn, bufs: Sequence[bytes], k, s, m, seed, extra=()):
    n = len(bufs)
    keep = [np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(0, np.uint8) for b in bufs]
    ptrs = (C.c_void_p * n)(*[a.ctypes.data if a.size else None for a in keep])
    lens = (C.c_size_t * n)(*[a.size for a in keep])
    oh = np.zeros(s, np.uint64)
    oc = np.zeros(s, np.uint32)
    st = _Stats()
    r = fn(ptrs, lens, n, k, s, m, seed, *extra, oh.ctypes.data, oc.ctypes.data, C.byref(st))
    if r < 0:
        raise ValueError(f"oracle: malformed record stream (code {r})")
    return OracleSketch(oh[:r].copy(), oc[:r].copy(), st.n_records, st.n_kept, st.total_len, st.n_kmers,
                        st.n_distinct, st.n_qualified, st.genome_size, st.coverage)


def sketch(bufs: Sequence[bytes], k=27, s=50000, m=1, seed=42) -> OracleSketch:
    """Exact-count sketch of the concatenation of `bufs` (mash sketch [-r -m])."""
    return _call_sketch(lib().orc_sketch, bufs, k, s, m, seed)


def sketch_stream(bufs: Sequence[bytes], k=27, s=50000, m=1, seed=42) -> OracleSketch:
    """Faithful single-thread streaming MinHashHeap (what AuriClass runs today)."""
    return _call_sketch(lib().orc_sketch_stream, bufs, k, s, m, seed)


def sketch_parallel(bufs: Sequence[bytes], k=27, s=50000, m=1, seed=42, threads=0) -> OracleSketch:
    return _call_sketch(lib().orc_sketch_parallel, bufs, k, s, m, seed, extra=(threads,))


def all_hashes(buf: bytes, k=27, seed=42) -> np.ndarray:
    a = np.frombuffer(buf, dtype=np.uint8)
    total = lib().orc_all_hashes(a.ctypes.data if a.size else None, a.size, k, seed, None, 0)
    if total < 0:
        raise ValueError("oracle: malformed record stream")
    out = np.zeros(max(total, 1), np.uint64)
    lib().orc_all_hashes(a.ctypes.data if a.size else None, a.size, k, seed, out.ctypes.data, total)
    return out[:total]


def murmur3_x64_128(key: bytes, seed=42) -> Tuple[int, int]:
    out = (C.c_uint64 * 2)()
    lib().orc_murmur3_x64_128(key, len(key), seed, C.byref(out))
    return int(out[0]), int(out[1])


def dist(ref: np.ndarray, qry: np.ndarray, sketch_size: int, k: int, len_ref: int, len_qry: int):
    ref = np.ascontiguousarray(ref, np.uint64)
    qry = np.ascontiguousarray(qry, np.uint64)
    d = _Dist()
    lib().orc_dist(ref.ctypes.data, ref.size, qry.ctypes.data, qry.size, sketch_size, k, len_ref, len_qry, C.byref(d))
    return d.distance, d.pvalue, int(d.common), int(d.denom)


def binomial_Q(k: int, p: float, n: int) -> float:
    return lib().orc_binomial_Q(k, p, n)


def bounds_text(k: int, prob: float) -> str:
    buf = C.create_string_buffer(1 << 16)
    n = lib().orc_bounds_text(k, prob, buf, len(buf))
    return buf.raw[:n].decode()


def num_threads() -> int:
    return lib().orc_num_threads()
