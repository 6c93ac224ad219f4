"""Host side of the B200 Mash path: stream FASTA/FASTQ(.gz) bytes into pinned
buffers, read the reference `.msh`, call libauriclass_cuda.so, and render the
results as the text `mash` would have printed.

Only the two host duties BASELINE.json's north_star leaves in Python live here
(file/gzip streaming and the Cap'n Proto reader); all arithmetic is CUDA.
Reference behaviour being mirrored: /root/reference/auriclass/classes.py:67-119
(`mash dist` + TSV parsing), :542-616 (`mash sketch -r -m`), :663-723
(`mash sketch`), SURVEY.md App. A.5/A.6 for Mash's output text.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass, field
from pathlib import Path
from typing import Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from ._lib import AC_MODE_FASTA, AC_MODE_READS, SketchInfo, check
from .msh import MshSketch, pack_references, read_msh

PathLike = Union[str, Path]
CHUNK_BYTES = 16 << 20  # pinned staging buffer size


def fmt_g(x: float) -> str:
    """C++ default ostream formatting of a double (= printf %g), as Mash prints."""
    return "%g" % x


@dataclass
class StreamStats:
    bytes_in: int = 0          # inflated bytes handed to the GPU
    read_s: float = 0.0        # file reads
    inflate_s: float = 0.0     # gzip inflate: thread-seconds inside zlib (host, timed separately per north_star)
    inflate_wall_s: float = 0.0  # wall time the feeder waited for inflated bytes (what parallel inflate shrinks)
    feed_s: float = 0.0        # time inside ac_sketch_feed (enqueue + occasional sync)
    end_s: float = 0.0         # ac_sketch_end
    total_s: float = 0.0


class GpuSketch:
    """Device-resident bottom-s sketch (replaces the temporary query `.msh`)."""

    def __init__(self, engine: "MashGPU", handle: int, info: SketchInfo, name: str, k: int, s: int):
        self._engine, self._h, self.name, self.k, self.s = engine, handle, name, k, s
        self.info = {f: getattr(info, f) for f, _ in SketchInfo._fields_}
        self.stats: Optional[StreamStats] = None

    @property
    def handle(self) -> int:
        if not self._h:
            raise ValueError("sketch has been freed")
        return self._h

    def __len__(self) -> int:
        return int(_lib.load().ac_sketch_size(self.handle))

    @property
    def genome_size(self) -> float:
        return float(self.info["genome_size"])

    @property
    def coverage(self) -> float:
        return float(self.info["coverage"])

    @property
    def length(self) -> int:
        return int(_lib.load().ac_sketch_length(self.handle))

    def hashes(self) -> Tuple[np.ndarray, np.ndarray]:
        n = len(self)
        h = np.zeros(n, np.uint64)
        c = np.zeros(n, np.uint32)
        check(_lib.load().ac_sketch_hashes(self.handle, h.ctypes.data, c.ctypes.data))
        return h, c

    def free(self) -> None:
        if self._h:
            _lib.load().ac_sketch_free(self._h)
            self._h = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class GpuReferences:
    """Reference sketches of one `.msh`, resident on the device."""

    def __init__(self, handle: int, sketch: MshSketch):
        self._h = handle
        self.names = [r.name for r in sketch.references]
        self.lengths = [r.length for r in sketch.references]
        self.kmer_size, self.sketch_size, self.hash_bits, self.seed = (
            sketch.kmer_size, sketch.sketch_size, sketch.hash_bits, sketch.seed)

    @property
    def handle(self) -> int:
        return self._h

    def free(self) -> None:
        if self._h:
            _lib.load().ac_refs_free(self._h)
            self._h = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


@dataclass
class DistRow:
    reference: str
    query: str
    distance: float
    pvalue: float
    common: int
    denom: int

    def line(self) -> str:
        # Mash: ref \t query \t dist \t p \t common/denom  (CommandDistance.cpp writeOutput)
        return f"{self.reference}\t{self.query}\t{fmt_g(self.distance)}\t{fmt_g(self.pvalue)}\t{self.common}/{self.denom}"


class _PinnedFeeder:
    """Copies host bytes of ONE file through the engine's two pinned staging buffers into ac_sketch_feed; the copy into
    buffer B overlaps the H2D transfer of buffer A."""

    def __init__(self, engine: "MashGPU", job: int, file_idx: int, st: StreamStats):
        self.eng, self.job, self.file_idx, self.st = engine, job, file_idx, st
        self.bufs = engine._staging()
        self.which = 0
        self.in_flight = 0
        self.fill = 0

    def _feed(self, n: int, eof: int) -> None:
        t0 = time.perf_counter()
        lib = self.eng.lib
        check(lib.ac_sketch_feed(self.job, self.bufs[self.which][0], n, self.file_idx, eof))
        self.st.bytes_in += n
        self.which ^= 1
        self.in_flight += 1
        if self.in_flight >= 2 and not eof:
            # both staging buffers are in flight: wait before overwriting the older one
            check(lib.ac_sketch_feed_wait(self.job))
            self.in_flight = 0
        self.st.feed_s += time.perf_counter() - t0

    def write(self, data) -> None:
        mv = memoryview(data)
        off = 0
        while off < len(mv):
            n = min(CHUNK_BYTES - self.fill, len(mv) - off)
            self.bufs[self.which][1][self.fill:self.fill + n] = mv[off:off + n]
            self.fill += n
            off += n
            if self.fill == CHUNK_BYTES:
                self._feed(self.fill, 0)
                self.fill = 0

    def close(self) -> None:
        """End of the file: the rest of the staging buffer goes out with the EOF mark."""
        self._feed(self.fill, 1)
        self.fill = 0
        check(self.eng.lib.ac_sketch_feed_wait(self.job))


class MashGPU:
    """One device context (ac_ctx).  Not thread safe; one per thread / rank."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.ac_create(device, C.byref(h)))
        self._ctx = h.value
        self.device = device
        self._pinned: List[int] = []
        self._refs_cache = {}
        self._sharders = {}
        self.inflate_threads: Optional[int] = None   # None: inflate.default_threads()

    # ---------------------------------------------------------------- plumbing
    @property
    def ctx(self) -> int:
        if not self._ctx:
            raise ValueError("context closed")
        return self._ctx

    def close(self) -> None:
        if self._ctx:
            for sh in self._sharders.values():
                sh.close()
            self._sharders.clear()
            for r in self._refs_cache.values():
                r.free()
            self._refs_cache.clear()
            for p in self._pinned:
                self.lib.ac_free_pinned(p)
            self._pinned = []
            self.lib.ac_destroy(self._ctx)
            self._ctx = 0

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self) -> int:
        return int(self.lib.ac_launch_count(self.ctx))

    def set_profile(self, on: bool) -> None:
        check(self.lib.ac_set_profile(self.ctx, 1 if on else 0))

    def last_job_times(self) -> dict:
        arr = (C.c_float * 4)()
        check(self.lib.ac_last_job_times(self.ctx, C.byref(arr)))
        return {"parse_ms": arr[0], "hash_ms": arr[1], "select_ms": arr[2], "total_ms": arr[3]}

    def int_peak(self) -> float:
        v = C.c_double()
        check(self.lib.ac_int_peak(self.ctx, C.byref(v)))
        return v.value

    def pipe_rates(self) -> dict:
        arr = (C.c_double * 3)()
        check(self.lib.ac_pipe_rates(self.ctx, C.byref(arr)))
        return {"imad": arr[0], "imad_wide": arr[1], "lop3": arr[2]}

    def alu_rates(self) -> dict:
        arr = (C.c_double * 5)()
        check(self.lib.ac_alu_rates(self.ctx, C.byref(arr)))
        return {"shf": arr[0], "prmt": arr[1], "iadd3": arr[2], "isetp_sel": arr[3], "lea": arr[4]}

    def _staging(self) -> List[Tuple[int, memoryview]]:
        if not self._pinned:
            for _ in range(2):
                p = self.lib.ac_alloc_pinned(CHUNK_BYTES)
                if not p:
                    raise MemoryError("ac_alloc_pinned failed: " + self.lib.ac_last_error().decode())
                self._pinned.append(p)
        return [(p, memoryview((C.c_ubyte * CHUNK_BYTES).from_address(p)).cast("B")) for p in self._pinned]

    # ---------------------------------------------------------------- sketching
    def _begin(self, k: int, s: int, seed: int, m: int, mode: int) -> int:
        job = C.c_void_p()
        check(self.lib.ac_sketch_begin(self.ctx, k, s, seed, m, mode, C.byref(job)))
        return job.value

    def _stream_file(self, job: int, path: PathLike, file_idx: int, st: StreamStats) -> None:
        """Duty (2) of the host: file bytes -> (inflate, inflate.py) -> pinned buffer -> ac_sketch_feed."""
        from . import inflate
        feeder = _PinnedFeeder(self, job, file_idx, st)
        for data in inflate.iter_inflated(path, st, self.inflate_threads):
            feeder.write(data)
        feeder.close()

    def _end(self, job: int, name: str, k: int, s: int, st: StreamStats) -> GpuSketch:
        sk = C.c_void_p()
        info = SketchInfo()
        t0 = time.perf_counter()
        check(self.lib.ac_sketch_end(job, C.byref(sk), C.byref(info)))
        st.end_s += time.perf_counter() - t0
        out = GpuSketch(self, sk.value, info, name, k, s)
        out.stats = st
        return out

    def sketch_reads(self, paths: Sequence[PathLike], k: int = 27, s: int = 50000, m: int = 3,
                     seed: int = 42) -> GpuSketch:
        """`mash sketch -r -m m -k k -s s paths...`: all files -> ONE sketch named after the first."""
        st = StreamStats()
        t0 = time.perf_counter()
        job = self._begin(k, s, seed, m, AC_MODE_READS)
        try:
            for i, p in enumerate(paths):
                self._stream_file(job, p, i, st)
        except BaseException:
            self.lib.ac_sketch_abort(job)
            raise
        sk = self._end(job, str(paths[0]), k, s, st)
        st.total_s = time.perf_counter() - t0
        return sk

    def sharder(self, group=None):
        """The exchange protocol object of this engine for a torch.distributed process group (made once per group)."""
        import torch.distributed as dist
        from .sharded import ShardedSketcher
        key = id(group)
        sh = self._sharders.get(key)
        if sh is None:
            sh = ShardedSketcher(self, dist.get_rank(group), dist.get_world_size(group), group)
            self._sharders[key] = sh
        return sh

    def sketch_reads_sharded(self, paths: Sequence[PathLike], k: int = 27, s: int = 50000, m: int = 3, seed: int = 42,
                             group=None, block_bytes: Optional[int] = None) -> GpuSketch:
        """`mash sketch -r -m m -k k -s s paths...` over the ranks of a torch.distributed process group (one rank per
        GPU; BASELINE.json configs[2]).  COLLECTIVE: every rank calls it with the same arguments and gets the same
        sketch.  The inflated streams are cut at record boundaries into read blocks (readblocks.py), block b goes to
        rank b mod P; each rank parses and hashes its blocks, then the ranks make the counts exact by one hash-range
        exchange (sharded.py).  k-mers never span records, so the result is the sketch of all the reads."""
        import torch.distributed as dist
        from . import inflate
        from .readblocks import DEFAULT_BLOCK_BYTES, ReadBlockCutter
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        sh = self.sharder(group)
        st = StreamStats()
        t0 = time.perf_counter()
        cutter = ReadBlockCutter(block_bytes or DEFAULT_BLOCK_BYTES)
        job = self._begin(k, s, seed, m, AC_MODE_READS)
        failure = None
        try:
            for i, p in enumerate(paths):
                cutter.start_file()
                feeder = _PinnedFeeder(self, job, i, st)
                for data in inflate.iter_inflated(p, st, self.inflate_threads):
                    for bi, piece in cutter.feed(data):
                        if bi % world == rank:
                            feeder.write(piece)
                for bi, piece in cutter.end_file():
                    if bi % world == rank:
                        feeder.write(piece)
                feeder.close()
        except BaseException as e:    # the other ranks are on their way into the exchange: tell them there, then raise
            failure = e
        sk = sh.finish(job, k, s, m, name=str(paths[0]), failure=failure)
        sk.stats = st
        st.total_s = time.perf_counter() - t0
        return sk

    def sketch_fasta(self, path: PathLike, k: int = 27, s: int = 50000, seed: int = 42) -> GpuSketch:
        """`mash sketch -k k -s s path`: one sketch per file, every k-mer counts (m = 1)."""
        st = StreamStats()
        t0 = time.perf_counter()
        job = self._begin(k, s, seed, 1, AC_MODE_FASTA)
        try:
            self._stream_file(job, path, 0, st)
        except BaseException:
            self.lib.ac_sketch_abort(job)
            raise
        sk = self._end(job, str(path), k, s, st)
        st.total_s = time.perf_counter() - t0
        from . import fastx
        fastx.record_size(path, sk.info["n_bases"])   # pyfastx `Fasta(path).size` for the FASTA genome-size QC
        return sk

    def sketch_bytes(self, bufs: Sequence[bytes], k: int = 27, s: int = 50000, m: int = 1, seed: int = 42,
                     mode: int = AC_MODE_READS, name: str = "bytes", chunk: Optional[int] = None) -> GpuSketch:
        """Sketch in-memory (inflated) buffers; `chunk` cuts the feeds at arbitrary byte counts (tests)."""
        st = StreamStats()
        job = self._begin(k, s, seed, m, mode)
        try:
            for i, b in enumerate(bufs):
                a = np.frombuffer(b, np.uint8)
                step = chunk or max(1, a.size)
                off = 0
                while True:
                    n = min(step, a.size - off)
                    last = off + n >= a.size
                    ptr = a[off:].ctypes.data if n else None
                    check(self.lib.ac_sketch_feed(job, ptr, n, i, 1 if last else 0))
                    off += n
                    if last:
                        break
            check(self.lib.ac_sketch_feed_wait(job))
        except BaseException:
            self.lib.ac_sketch_abort(job)
            raise
        return self._end(job, name, k, s, st)

    def sketch_from_hashes(self, hashes: np.ndarray, k: int, s: int, length: int, name: str = "query") -> GpuSketch:
        h = np.ascontiguousarray(hashes, np.uint64)
        sk = C.c_void_p()
        check(self.lib.ac_sketch_from_hashes(self.ctx, h.ctypes.data, h.size, k, s, length, C.byref(sk)))
        info = SketchInfo()
        info.n_hashes = h.size
        return GpuSketch(self, sk.value, info, name, k, s)

    # ---------------------------------------------------------------- references + dist
    def load_references(self, msh: Union[PathLike, MshSketch]) -> GpuReferences:
        """Duty (1) of the host: parse the reference `.msh` (msh.py) and upload it once."""
        key = None
        if not isinstance(msh, MshSketch):
            key = str(Path(msh).resolve())
            if key in self._refs_cache:
                return self._refs_cache[key]
            msh = read_msh(msh)
        hashes, offsets, lengths = pack_references(msh.references)
        h = C.c_void_p()
        check(self.lib.ac_refs_load(self.ctx, hashes.ctypes.data if hashes.size else None, offsets.ctypes.data,
                                    lengths.ctypes.data if lengths.size else None, len(msh.references),
                                    msh.kmer_size, msh.sketch_size, msh.hash_bits, C.byref(h)))
        refs = GpuReferences(h.value, msh)
        if key:
            self._refs_cache[key] = refs
        return refs

    def dist(self, refs: GpuReferences, queries: Union[GpuSketch, Sequence[GpuSketch]]) -> List[DistRow]:
        """`mash dist REF QUERY...`: rows ordered query-major, references in file order."""
        qs = [queries] if isinstance(queries, GpuSketch) else list(queries)
        nr, nq = len(refs.names), len(qs)
        d = np.zeros(nq * nr, np.float64)
        p = np.zeros(nq * nr, np.float64)
        cm = np.zeros(nq * nr, np.uint64)
        dn = np.zeros(nq * nr, np.uint64)
        arr = (C.c_void_p * max(nq, 1))(*[q.handle for q in qs])
        check(self.lib.ac_dist_batch(self.ctx, refs.handle, arr, nq, d.ctypes.data, p.ctypes.data,
                                     cm.ctypes.data, dn.ctypes.data))
        rows = []
        for qi, q in enumerate(qs):
            for ri, name in enumerate(refs.names):
                o = qi * nr + ri
                rows.append(DistRow(name, q.name, float(d[o]), float(p[o]), int(cm[o]), int(dn[o])))
        return rows


    def load_sketch_set(self, sketches: Sequence[np.ndarray], lengths: Sequence[int], k: int, s: int,
                        names: Optional[Sequence[str]] = None) -> GpuReferences:
        """Upload many finished sketches (ascending hash arrays) as one device-resident set: the operand of
        `dist_pairs` (all-vs-all `mash dist`, BASELINE config 5) or a reference set built in memory."""
        from .msh import MshReference
        refs = [MshReference(names[i] if names else f"sketch_{i}", "", int(lengths[i]),
                             np.ascontiguousarray(h, np.uint64)) for i, h in enumerate(sketches)]
        return self.load_references(MshSketch(kmer_size=k, sketch_size=s, hash_bits=64 if k > 16 else 32, references=refs))

    def dist_pairs(self, a: GpuReferences, i0: int, ni: int, b: GpuReferences, j0: int, nj: int):
        """One tile of `mash dist A B`: sketches a[i0:i0+ni] x b[j0:j0+nj] -> (distance, pvalue, common, denom),
        each an [ni, nj] array."""
        d = np.zeros((ni, nj), np.float64)
        p = np.zeros((ni, nj), np.float64)
        cm = np.zeros((ni, nj), np.uint64)
        dn = np.zeros((ni, nj), np.uint64)
        check(self.lib.ac_dist_pairs(self.ctx, a.handle, i0, ni, b.handle, j0, nj, d.ctypes.data, p.ctypes.data,
                                     cm.ctypes.data, dn.ctypes.data))
        return d, p, cm, dn


def dist_stdout(rows: Iterable[DistRow]) -> str:
    """stdout of `mash dist` for these rows."""
    return "".join(r.line() + "\n" for r in rows)


def sketch_reads_stderr(sk: GpuSketch, out_path: PathLike) -> str:
    """stderr of `mash sketch -r -m`: the two estimate lines + `Writing to`."""
    return (f"Estimated genome size: {fmt_g(sk.genome_size)}\n"
            f"Estimated coverage:    {fmt_g(sk.coverage)}\n"
            f"Writing to {out_path}...\n")


def sketch_fasta_stderr(paths: Sequence[PathLike], out_path: PathLike) -> str:
    return "".join(f"Sketching {p}...\n" for p in paths) + f"Writing to {out_path}...\n"


NO_RECORDS_STDERR = 'ERROR: Did not find fasta records in "{}".\n'
