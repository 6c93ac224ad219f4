"""ctypes binding of libauriclass_cuda.so (include/auriclass_cuda.h).

This is the reference-side stub a maintainer adds where AuriClass used to call
`subprocess.run(["mash", ...])` (/root/reference/auriclass/classes.py:100-104,
:592-596, :709-713; general.py:198-205).  There is no CPU fallback: if the
shared library is missing or no sm_100 device is visible, every entry point
raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("AURICLASS_CUDA_LIB", _HERE / "csrc" / "libauriclass_cuda.so"))

(AC_OK, AC_E_PARAM, AC_E_CUDA, AC_E_NOMEM, AC_E_NO_RECORDS, AC_E_FORMAT, AC_E_PARAM_MISMATCH, AC_E_STATE, AC_E_INTERNAL,
 AC_E_PEER) = range(10)
AC_MODE_READS, AC_MODE_FASTA = 0, 1


class SketchInfo(C.Structure):
    _fields_ = [
        ("n_hashes", C.c_uint64), ("n_records", C.c_uint64), ("n_kept", C.c_uint64),
        ("total_len", C.c_uint64), ("n_kmers", C.c_uint64), ("n_candidates", C.c_uint64),
        ("max_hash", C.c_uint64), ("sum_counts", C.c_uint64), ("final_tau", C.c_uint64),
        ("genome_size", C.c_double), ("coverage", C.c_double),
        ("hash_bits", C.c_int32), ("passes", C.c_int32), ("n_bases", C.c_uint64),
    ]


class AcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libauriclass_cuda: {msg} (code {code})")
        self.code = code
        self.msg = msg


class NoRecordsError(AcError):
    """Mash's `ERROR: Did not find fasta records in ...` (AC_E_NO_RECORDS)."""


class PeerError(AcError):
    """Sharded sketch: another rank of the exchange failed or did not answer (AC_E_PEER)."""


# every symbol include/auriclass_cuda.h declares: name -> (restype, argtypes)
_vp, _u64p, _u32p, _dp = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_double)
SIGNATURES = {
    "ac_abi_version": (C.c_int, []),
    "ac_probe": (C.c_int, []),
    "ac_device_count": (C.c_int, []),
    "ac_last_error": (C.c_char_p, []),
    "ac_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "ac_destroy": (None, [_vp]),
    "ac_sync": (C.c_int, [_vp]),
    "ac_set_stream": (C.c_int, [_vp, _vp, C.c_int]),
    "ac_alloc_pinned": (_vp, [C.c_size_t]),
    "ac_free_pinned": (None, [_vp]),
    "ac_sketch_begin": (C.c_int, [_vp, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.POINTER(_vp)]),
    "ac_sketch_feed": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, C.c_int]),
    "ac_sketch_feed_device": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, C.c_int]),
    "ac_sketch_feed_wait": (C.c_int, [_vp]),
    "ac_sketch_end": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(SketchInfo)]),
    "ac_sketch_abort": (None, [_vp]),
    "ac_sketch_end_local": (C.c_int, [_vp, _u64p, _u64p, C.POINTER(C.c_int), C.POINTER(SketchInfo)]),
    "ac_job_table_ranges": (C.c_int, [_vp, C.c_uint64, C.c_int, _u64p]),
    "ac_job_table_export": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp, _vp]),
    "ac_merge_counts": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_int, _u64p]),
    "ac_sketch_from_table": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(SketchInfo)]),
    "ac_xchg_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(_vp)]),
    "ac_xchg_ipc_handle": (C.c_int, [_vp, _vp]),
    "ac_xchg_open_peers": (C.c_int, [_vp, _vp]),
    "ac_xchg_set_peer_local": (C.c_int, [_vp, C.c_int, _vp]),
    "ac_xchg_destroy": (None, [_vp]),
    "ac_sketch_exchange_begin": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ac_sketch_exchange_merge": (C.c_int, [_vp]),
    "ac_sketch_exchange_end": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(SketchInfo), C.POINTER(C.c_uint64 * 4)]),
    "ac_sketch_from_hashes": (C.c_int, [_vp, _vp, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.POINTER(_vp)]),
    "ac_sketch_size": (C.c_uint64, [_vp]),
    "ac_sketch_hashes": (C.c_int, [_vp, _vp, _vp]),
    "ac_sketch_length": (C.c_uint64, [_vp]),
    "ac_sketch_free": (None, [_vp]),
    "ac_refs_load": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "ac_refs_count": (C.c_int, [_vp]),
    "ac_refs_free": (None, [_vp]),
    "ac_dist": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ac_dist_batch": (C.c_int, [_vp, _vp, C.POINTER(_vp), C.c_int, _vp, _vp, _vp, _vp]),
    "ac_dist_pairs": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "ac_launch_count": (C.c_uint64, [_vp]),
    "ac_timer_start": (C.c_int, [_vp]),
    "ac_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "ac_set_profile": (C.c_int, [_vp, C.c_int]),
    "ac_last_job_times": (C.c_int, [_vp, C.POINTER(C.c_float * 4)]),
    "ac_int_peak": (C.c_int, [_vp, _dp]),
    "ac_pipe_rates": (C.c_int, [_vp, C.POINTER(C.c_double * 3)]),
    "ac_alu_rates": (C.c_int, [_vp, C.POINTER(C.c_double * 5)]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (no CUDA call is made).  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  auriclass_b200 has no CPU fallback.")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc == AC_OK:
        return
    msg = load().ac_last_error().decode("utf-8", "replace")
    if rc == AC_E_NO_RECORDS:
        raise NoRecordsError(rc, msg)
    if rc == AC_E_PEER:
        raise PeerError(rc, msg)
    raise AcError(rc, msg)
