"""Read-block sharded `mash sketch -r -m` over the GPUs of one box (BASELINE.json configs[2]).

Every rank parses and hashes its own read blocks (ac_sketch_begin/feed on its GPU) and ends with an
exact local table {hash <= tau_r : local count}.  The sketch Mash would have produced from all the
reads is bottom-s of {h : sum over ranks of count_r(h) >= m} (the result is
order independent), so one exchange step makes the counts exact:

  1. tau_cut = min_r tau_r                       (all-reduce MIN, 8 bytes)
     every rank holds exact local counts of everything <= tau_cut, and the rank that attained the
     minimum proves >= s qualified hashes exist below it (or tau_cut is the whole hash space).
  2. hash-range exchange                         (all-to-all: counts, then hashes and counts)
     [0, tau_cut] is cut into P equal ranges; range r is owned by rank r.  Hashes are uniform, so
     the ranges balance.  Volume: the local tables, ~1e5..1e6 entries of 12 bytes per rank.
  3. owner merge: sum counts of equal hashes, keep >= m, sort   (ac_merge_counts: CUDA radix sort + run sums)
  4. all-gather of the owners' first <= s qualified entries; the ordered concatenation's first s
     entries are the sketch                      (ac_sketch_from_table: K4 estimates)

The collectives go through torch.distributed (NCCL over NVLink on the GPUs; gloo in the CPU tests of
this host logic).  There is no fused compute+collective kernel: the exchange moves about a megabyte
once per isolate, after the last hash kernel has finished, so there is nothing to overlap.

The arithmetic sits behind a small backend interface so the same exchange logic runs in
tests/test_sharded_gloo.py with tables made by the CPU oracle; the product backend is
`CudaTableBackend` (libauriclass_cuda.so) and nothing else.

Replaces: `mash sketch -r -m M` at /root/reference/auriclass/classes.py:576-596 for inputs large
enough to shard (the reference itself is single process, docs/running_analysis.md:47-60).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import SketchInfo, check

_SIGN = -(1 << 63)  # u64 -> order-preserving i64: x ^ 0x8000_0000_0000_0000


def u64_to_ordered_i64(x: int) -> int:
    x &= (1 << 64) - 1
    y = x ^ (1 << 63)
    return y - (1 << 64) if y >= (1 << 63) else y


def ordered_i64_to_u64(y: int) -> int:
    return (y & ((1 << 64) - 1)) ^ (1 << 63)


def range_bounds(tau_cut: int, n_ranges: int) -> List[int]:
    """b[r] = first hash owned by range r over [0, tau_cut]: owner(x) = floor(x * P / (tau_cut + 1))."""
    span = tau_cut + 1
    return [(span * r + n_ranges - 1) // n_ranges for r in range(n_ranges)]


class PeerFailure(_lib.PeerError):
    """Another rank of the group failed before the exchange (bad input in its read blocks, CUDA error, ...): raised
    on every rank at the first collective, so nobody is left waiting."""

    def __init__(self, n_failed: int):
        super().__init__(_lib.AC_E_PEER, f"sharded sketch aborted: {n_failed} rank(s) of the group failed before the exchange")


class PeerWindow:
    """This rank's window of the NVLink peer-memory exchange (csrc/ac_xchg.cu, include/auriclass_cuda.h `ac_xchg_*`):
    the hash-range exchange as kernels that store straight into the other GPUs' memory -- no collective library call
    and two small host reads per isolate."""

    def __init__(self, engine, rank: int, world: int, sketch_cap: int, inbox_entries: int = 0):
        self.eng, self.lib, self.rank, self.world, self.sketch_cap = engine, engine.lib, rank, world, sketch_cap
        h = C.c_void_p()
        check(self.lib.ac_xchg_create(engine.ctx, rank, world, sketch_cap, inbox_entries, C.byref(h)))
        self._h = h.value

    def handle_bytes(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(self.lib.ac_xchg_ipc_handle(self._h, buf))
        return buf.raw

    def open_peers(self, handles) -> None:
        blob = b"".join(bytes(h) for h in handles)
        assert len(blob) == 64 * self.world
        check(self.lib.ac_xchg_open_peers(self._h, blob))

    def set_peer_local(self, r: int, other: "PeerWindow") -> None:
        check(self.lib.ac_xchg_set_peer_local(self._h, r, other._h))

    def begin(self, job, k: int, s: int, m: int, failed: bool = False) -> None:
        check(self.lib.ac_sketch_exchange_begin(self.eng.ctx, job or None, self._h, k, s, m, 1 if failed else 0))

    def merge(self) -> None:
        check(self.lib.ac_sketch_exchange_merge(self._h))

    def end(self, name: str, k: int, s: int):
        from .mash_gpu import GpuSketch
        sk = C.c_void_p()
        info = SketchInfo()
        stats = (C.c_uint64 * 4)()
        check(self.lib.ac_sketch_exchange_end(self._h, C.byref(sk), C.byref(info), C.byref(stats)))
        out = GpuSketch(self.eng, sk.value, info, name, k, s)
        self.last_stats = {"tau_cut": int(stats[0]), "received": int(stats[1]), "qualified_total": int(stats[2])}
        return out

    def close(self) -> None:
        if self._h:
            self.lib.ac_xchg_destroy(self._h)
            self._h = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class LocalTable:
    """A rank's exact (hash, count) table of everything <= tau, ascending; bit patterns of u64 / u32
    carried in int64 / int32 tensors (torch collectives have no unsigned 64-bit type)."""
    hashes: torch.Tensor   # int64, device of the backend
    counts: torch.Tensor   # int32
    tau: int
    counters: Tuple[int, ...] = (0, 0, 0, 0, 0)  # n_records, n_kept, total_len, n_kmers, n_candidates


class CudaTableBackend:
    """libauriclass_cuda.so on this rank's GPU."""

    def __init__(self, engine):
        self.eng = engine
        self.lib = engine.lib
        self.device = torch.device("cuda", engine.device)
        self.stream_shared = False

    def share_torch_stream(self) -> None:
        """Run the library's kernels on torch's current stream: the collectives torch.distributed queues on that
        stream are then ordered against them on the device and no host synchronisation is needed in between."""
        check(self.lib.ac_set_stream(self.eng.ctx, torch.cuda.current_stream(self.device).cuda_stream, 1))
        self.stream_shared = True

    def unshare_stream(self) -> None:
        if self.stream_shared:
            check(self.lib.ac_set_stream(self.eng.ctx, None, 0))
            self.stream_shared = False

    def local_table(self, job: int, tau_cut_fn, n_ranges: int):
        """End the local pass; returns (send splits, LocalTable restricted to <= tau_cut)."""
        n = C.c_uint64()
        tau = C.c_uint64()
        spec = C.c_int()
        info = SketchInfo()
        try:
            try:
                check(self.lib.ac_sketch_end_local(job, C.byref(n), C.byref(tau), C.byref(spec), C.byref(info)))
            except BaseException:
                try:
                    tau_cut_fn((1 << 64) - 1, False)   # the others are waiting in the all-reduce: tell them
                except PeerFailure:
                    pass
                raise
            tau_cut = tau_cut_fn(int(tau.value))
            begins = (C.c_uint64 * (n_ranges + 1))()
            check(self.lib.ac_job_table_ranges(job, tau_cut, n_ranges, begins))
            b = [int(x) for x in begins]
            n_send = b[n_ranges]
            h = torch.empty(max(n_send, 1), dtype=torch.int64, device=self.device)
            c = torch.empty(max(n_send, 1), dtype=torch.int32, device=self.device)
            check(self.lib.ac_job_table_export(job, 0, n_send, h.data_ptr(), c.data_ptr()))
        finally:
            self.lib.ac_sketch_abort(job)  # frees the job; the exported copy is ours
        splits = [b[r + 1] - b[r] for r in range(n_ranges)]
        counters = (info.n_records, info.n_kept, info.total_len, info.n_kmers, info.n_candidates)
        return splits, LocalTable(h[:n_send], c[:n_send], tau_cut, counters)

    def merge_counts(self, hashes: torch.Tensor, counts: torch.Tensor, min_copies: int) -> int:
        """In place: sum equal hashes, keep count >= m, ascending.  Returns the number kept."""
        n_out = C.c_uint64()
        n = hashes.numel()
        if n == 0:
            return 0
        if not self.stream_shared:
            torch.cuda.current_stream(self.device).synchronize()  # NCCL wrote these on torch's stream; the library has its own
        check(self.lib.ac_merge_counts(self.eng.ctx, hashes.data_ptr(), counts.data_ptr(), n, min_copies, C.byref(n_out)))
        return int(n_out.value)

    def build_sketch(self, hashes: torch.Tensor, counts: torch.Tensor, k: int, s: int, name: str):
        from .mash_gpu import GpuSketch
        sk = C.c_void_p()
        info = SketchInfo()
        if not self.stream_shared:
            torch.cuda.current_stream(self.device).synchronize()
        check(self.lib.ac_sketch_from_table(self.eng.ctx, hashes.data_ptr(), counts.data_ptr(), hashes.numel(), k, s,
                                            C.byref(sk), C.byref(info)))
        return GpuSketch(self.eng, sk.value, info, name, k, s)


class ShardedSketcher:
    """The exchange protocol (steps 1-4 above) over a torch.distributed process group."""

    def __init__(self, engine, rank: int, world: int, group: Optional[dist.ProcessGroup] = None, backend=None):
        self.rank, self.world, self.group = rank, world, group
        self.engine = engine
        self.backend = backend if backend is not None else CudaTableBackend(engine)
        self.device = self.backend.device
        self.last_exchange = {}
        self._mark = None   # optional callable(name): phase timing hook (tools/time_sharded.py)
        # Transport: kernels over NVLink peer memory when every rank sits on the same box and can open the others'
        # windows (the product path on an NVSwitch box); else torch.distributed collectives (NCCL between boxes, gloo in
        # the CPU tests of this host logic).  AC_XCHG=collectives forces the latter.
        import os
        self.peer: Optional[PeerWindow] = None
        self._peer_ok = (isinstance(self.backend, CudaTableBackend) and world <= 16
                         and os.environ.get("AC_XCHG", "peer") != "collectives")
        if self._peer_ok:
            self._peer_window(65536)     # collective, like the construction of a sharder itself
        self._set_transport()

    TORCH_TRANSPORT = ("torch.distributed collectives: all-reduce MIN, all-to-all of (hash, count) ranges, "
                       "all-gather of the owners' prefixes")
    PEER_TRANSPORT = ("kernels over NVLink peer memory (CUDA IPC windows): tau post/min, scatter by owner into the owners' "
                      "inboxes, owner merge, prefix push to every rank; no collective library call")

    def _set_transport(self) -> None:
        self.transport = self.PEER_TRANSPORT if self._peer_ok else self.TORCH_TRANSPORT
        if not self._peer_ok and hasattr(self.backend, "share_torch_stream") and not self.backend.stream_shared:
            try:
                self.backend.share_torch_stream()
            except _lib.AcError:
                pass    # a job is in flight on the library's own stream: the backend then synchronises the streams instead

    def _peer_window(self, s: int) -> Optional[PeerWindow]:
        """(Re)build the windows, collectively, when there is none yet or the sketch size outgrew it.  All ranks must
        succeed, else all fall back to the collectives."""
        if self.peer is not None and self.peer.sketch_cap >= s:
            return self.peer
        import os
        import socket
        if self.peer is not None:
            self.peer.close()
            self.peer = None
        win, handle, err = None, b"", ""
        try:
            win = PeerWindow(self.engine, self.rank, self.world, max(int(s), 65536), int(os.environ.get("AC_XCHG_INBOX", "0")))
            handle = win.handle_bytes()
        except Exception as e:   # noqa: BLE001 -- reported to the group below
            err = repr(e)
        infos = [None] * self.world
        dist.all_gather_object(infos, (socket.gethostname(), handle, err), group=self.group)
        ok = all(i[2] == "" for i in infos) and len({i[0] for i in infos}) == 1
        if ok:
            try:
                win.open_peers([i[1] for i in infos])
            except Exception as e:   # noqa: BLE001
                err = repr(e)
        oks = [None] * self.world
        dist.all_gather_object(oks, ok and err == "", group=self.group)
        if all(oks):
            self.peer = win
            return win
        if win is not None:
            win.close()
        if self.rank == 0:
            import logging
            logging.warning("auriclass_b200: peer-memory exchange unavailable (%s); using torch.distributed collectives",
                            "; ".join(i[2] for i in infos if i[2]) or err or "ranks on different hosts")
        self._peer_ok = False
        self._set_transport()
        return None

    def close(self) -> None:
        if self.peer is not None:
            self.peer.close()
            self.peer = None
        if hasattr(self.backend, "unshare_stream"):
            self.backend.unshare_stream()

    # -- step 1 (also the point where the ranks agree that everybody got this far)
    def _tau_cut(self, tau_local: int, ok: bool = True) -> int:
        t = torch.tensor([u64_to_ordered_i64(tau_local), 0 if ok else -1], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        tau_cut, worst = t.tolist()
        if worst < 0:
            f = torch.tensor([0 if ok else 1], dtype=torch.int64, device=self.device)
            dist.all_reduce(f, op=dist.ReduceOp.SUM, group=self.group)
            raise PeerFailure(int(f.item()))
        return ordered_i64_to_u64(int(tau_cut))

    def finish(self, job, k: int, s: int, m: int, name: str = "reads", failure: Optional[BaseException] = None):
        """Collective: every rank calls it with its own job; every rank returns the same sketch.  `failure`: this rank
        could not even feed its reads (the job is aborted, the other ranks are told, the exception is re-raised)."""
        P = self.world
        mark = self._mark or (lambda name: None)
        win = self._peer_window(s) if self._peer_ok else None
        if win is not None:
            if failure is not None and job:
                _lib.load().ac_sketch_abort(job)
                job = None
            try:
                win.begin(job, k, s, m, failed=failure is not None)
                mark("xchg_begin")
                win.merge()
                mark("xchg_merge")
                sk = win.end(name, k, s)
                mark("xchg_end")
            except BaseException:
                if failure is not None:
                    raise failure
                raise
            self.last_exchange = dict(win.last_stats)
            return sk
        if failure is not None:
            if job:
                _lib.load().ac_sketch_abort(job)
            try:
                self._tau_cut((1 << 64) - 1, False)
            except PeerFailure:
                pass
            raise failure
        splits, tab = self.backend.local_table(job, self._tau_cut, P)
        mark("local_table")
        # -- step 2: sizes, then payload
        send_n = torch.tensor(splits, dtype=torch.int64, device=self.device)
        recv_n = torch.empty(P, dtype=torch.int64, device=self.device)
        dist.all_to_all_single(recv_n, send_n, group=self.group)
        rsplits = [int(x) for x in recv_n.tolist()]
        n_recv = sum(rsplits)
        rh = torch.empty(max(n_recv, 1), dtype=torch.int64, device=self.device)[:n_recv]
        rc = torch.empty(max(n_recv, 1), dtype=torch.int32, device=self.device)[:n_recv]
        dist.all_to_all_single(rh, tab.hashes, rsplits, splits, group=self.group)
        dist.all_to_all_single(rc, tab.counts, rsplits, splits, group=self.group)
        mark("all_to_all")
        # -- step 3: owner merge (exact counts of this rank's hash range)
        n_q = self.backend.merge_counts(rh, rc, m)
        n_keep = min(n_q, s)
        mark("merge_counts")
        # -- step 4: ONE all-gather of [header | <= s hashes | <= s counts] per owner (fixed size: no size exchange)
        HDR = 8
        pad = max(int(s), 1)
        words = HDR + pad + (pad + 1) // 2
        buf = torch.zeros(words, dtype=torch.int64, device=self.device)
        buf[:HDR] = torch.tensor([n_keep, n_q, *tab.counters, 0], dtype=torch.int64)
        if n_keep:
            buf[HDR:HDR + n_keep] = rh[:n_keep]
            buf[HDR + pad:].view(torch.int32)[:n_keep] = rc[:n_keep]
        gbuf = torch.empty(P * words, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(gbuf, buf, group=self.group)
        gbuf = gbuf.view(P, words)
        metas = gbuf[:, :HDR].tolist()
        mark("all_gather")
        keeps = [int(r[0]) for r in metas]
        # ordered concatenation (range r < range r+1), first s entries
        parts_h, parts_c, left = [], [], s
        for r in range(P):
            take = min(keeps[r], left)
            if take:
                parts_h.append(gbuf[r, HDR:HDR + take])
                parts_c.append(gbuf[r, HDR + pad:].view(torch.int32)[:take])
                left -= take
        if parts_h:
            fh, fc = torch.cat(parts_h).contiguous(), torch.cat(parts_c).contiguous()
        else:
            fh = torch.empty(0, dtype=torch.int64, device=self.device)
            fc = torch.empty(0, dtype=torch.int32, device=self.device)
        total_q = sum(int(r[1]) for r in metas)
        self.last_exchange = {"tau_cut": tab.tau, "sent": sum(splits), "received": n_recv, "qualified_owned": n_q,
                              "qualified_total": total_q}
        if fh.numel() == 0:
            raise _lib.NoRecordsError(_lib.AC_E_NO_RECORDS, "did not find sequence records (no hash reached the minimum copy number on any rank)")
        sk = self.backend.build_sketch(fh, fc, k, s, name)
        # whole-job counters (sums over ranks)
        for i, key in enumerate(("n_records", "n_kept", "total_len", "n_kmers", "n_candidates")):
            sk.info[key] = sum(int(r[2 + i]) for r in metas)
        sk.info["final_tau"] = tab.tau
        return sk
