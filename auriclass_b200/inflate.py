"""Host duty (2) of BASELINE.json's north_star: open FASTA/FASTQ(.gz) files and stream their *inflated* bytes
towards the pinned staging buffers -- with the inflate itself off the critical path where the container
format allows it (the reference leaves inflate to `mash`, i.e. zlib on one
thread: /root/reference/docs/running_analysis.md:47-60).

* plain file                -> read in chunks
* BGZF (bgzip, the format sequencing pipelines emit): every <= 64 KiB block carries its compressed size in
  a `BC` extra field, so the block table is found without inflating and the blocks are inflated
  independently on a pool of host threads (zlib releases the GIL), delivered in file order
* multi-member gzip (`cat a.gz b.gz`, pigz -i ...) without size fields, and ordinary single-member gzip:
  inherently sequential -- inflated on ONE background thread, so the inflate of chunk i+1 overlaps the
  H2D copy and the kernels of chunk i

Inflate time is accounted separately (`StreamStats.inflate_s` = thread-seconds spent in zlib,
`inflate_wall_s` = wall time the consumer waited for inflated bytes).
"""
from __future__ import annotations

import os
import queue
import struct
import threading
import time
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Iterator, List, Optional, Tuple, Union

PathLike = Union[str, Path]
READ_BYTES = 4 << 20
BGZF_BATCH = 64            # blocks (<= 4 MiB inflated) per pool task


def default_threads() -> int:
    return max(1, min(16, (os.cpu_count() or 2) - 1))


def _bgzf_block_size(hdr: bytes) -> Optional[int]:
    """Total size of the BGZF block whose header starts `hdr` (>= 18 bytes), or None if it is not BGZF."""
    if len(hdr) < 18 or hdr[:4] != b"\x1f\x8b\x08\x04":
        return None
    xlen = struct.unpack_from("<H", hdr, 10)[0]
    pos, end = 12, 12 + xlen
    while pos + 4 <= min(end, len(hdr)):
        si1, si2, slen = hdr[pos], hdr[pos + 1], struct.unpack_from("<H", hdr, pos + 2)[0]
        if si1 == 66 and si2 == 67 and slen == 2 and pos + 6 <= len(hdr):
            return struct.unpack_from("<H", hdr, pos + 4)[0] + 1
        pos += 4 + slen
    return None


def bgzf_block_table(buf: bytes) -> Optional[List[Tuple[int, int, int]]]:
    """[(offset, total block size, header length)] of every block if `buf` is a well-formed BGZF stream."""
    out, off, n = [], 0, len(buf)
    while off < n:
        size = _bgzf_block_size(buf[off:off + 64])
        if size is None or off + size > n or size < 26:
            return None
        xlen = struct.unpack_from("<H", buf, off + 10)[0]
        out.append((off, size, 12 + xlen))
        off += size
    return out


def _inflate_blocks(buf: bytes, blocks: List[Tuple[int, int, int]]) -> Tuple[bytes, float]:
    t0 = time.perf_counter()
    parts = []
    for off, size, hlen in blocks:
        raw = buf[off + hlen: off + size - 8]                 # deflate payload; trailer = CRC32 + ISIZE
        data = zlib.decompress(raw, wbits=-15)
        if (zlib.crc32(data) & 0xFFFFFFFF) != struct.unpack_from("<I", buf, off + size - 8)[0]:
            raise zlib.error(f"BGZF block at offset {off}: CRC mismatch")
        parts.append(data)
    return b"".join(parts), time.perf_counter() - t0


BGZF_WINDOW = 16 << 20     # compressed bytes read per step of the BGZF path (memory stays flat: ~2 windows + look-ahead)


def _iter_bgzf(f, head: bytes, threads: int, add):
    """BGZF path: the file is read in windows, each window's complete blocks are inflated on the pool in file order.
    Generator; its return value is None at a clean end of file, else the file offset where the block chain stopped
    being well-formed BGZF (the caller continues from there with the sequential inflater: a BGZF file is a
    multi-member gzip file, and so is whatever follows a block boundary)."""
    pos = 0                   # file offset of buf[0]
    buf = head
    eof = False
    with ThreadPoolExecutor(max_workers=threads) as pool:
        window: "queue.Queue" = queue.Queue()

        def drain(limit: int):
            while window.qsize() > limit:
                t0 = time.perf_counter()
                data, dt = window.get().result()
                add("inflate_wall_s", time.perf_counter() - t0)
                add("inflate_s", dt)
                if data:
                    yield data

        while True:
            if not eof and len(buf) < BGZF_WINDOW:
                t0 = time.perf_counter()
                more = f.read(BGZF_WINDOW)
                add("read_s", time.perf_counter() - t0)
                eof = not more
                buf += more
            # complete blocks at the head of buf
            table, off, n, bad = [], 0, len(buf), False
            while off < n:
                size = _bgzf_block_size(buf[off:off + 64])
                if size is None or size < 26:
                    bad = n - off >= 64 or eof          # a short tail of an unfinished read is not a verdict yet
                    break
                if off + size > n:
                    bad = eof                             # truncated last block
                    break
                xlen = struct.unpack_from("<H", buf, off + 10)[0]
                table.append((off, size, 12 + xlen))
                off += size
            for i in range(0, len(table), BGZF_BATCH):
                window.put(pool.submit(_inflate_blocks, buf, table[i:i + BGZF_BATCH]))
                yield from drain(2 * threads)             # bounded look-ahead
            if bad:
                yield from drain(0)
                return pos + off
            buf = buf[off:]
            pos += off
            if eof and not buf:
                yield from drain(0)
                return None


def iter_inflated(path: PathLike, stats=None, threads: Optional[int] = None) -> Iterator[bytes]:
    """Inflated bytes of `path`, in order, as a sequence of chunks."""
    threads = threads or default_threads()

    def add(field: str, dt: float) -> None:
        if stats is not None:
            setattr(stats, field, getattr(stats, field, 0.0) + dt)

    with open(path, "rb") as f:
        t0 = time.perf_counter()
        head = f.read(64)
        add("read_s", time.perf_counter() - t0)
        if head[:2] != b"\x1f\x8b":                           # ---- plain text
            if head:
                yield head
            while True:
                t0 = time.perf_counter()
                chunk = f.read(READ_BYTES)
                add("read_s", time.perf_counter() - t0)
                if not chunk:
                    return
                yield chunk
        if _bgzf_block_size(head) is not None:                # ---- BGZF: parallel block inflate
            resume = yield from _iter_bgzf(f, head, threads, add)
            if resume is None:
                return
            f.seek(resume)                                    # malformed block chain: sequential inflate from there
            resumed = True
        else:
            f.seek(0)
            resumed = False

        # ---- sequential gzip (single or multi member): one background thread, bounded queue
        q: "queue.Queue" = queue.Queue(maxsize=4)
        stop = threading.Event()

        def producer() -> None:
            try:
                dec = None
                at_boundary = True     # between members: a gzip header must follow, anything else is trailing garbage
                n_members = 0
                src = b""
                while not stop.is_set():
                    t0 = time.perf_counter()
                    more = f.read(1 << 20)
                    add("read_s", time.perf_counter() - t0)
                    src += more
                    if not more:
                        if not at_boundary or (src and src == b"\x1f\x8b"[:len(src)]):
                            raise zlib.error("gzip stream ends inside a member")
                        break                                  # (a lone trailing byte that is not a header start: garbage)
                    while src:
                        t0 = time.perf_counter()
                        if at_boundary:
                            # what follows a finished member: another member (concatenated members are legal) or
                            # padding / garbage, which zlib's gzread -- what mash reads its inputs with -- ignores
                            if len(src) < 2:
                                break                          # decide once two bytes are there
                            if src[:2] != b"\x1f\x8b":
                                if n_members == 0 and not resumed:
                                    raise zlib.error("not a gzip stream")
                                src = b""
                                stop.set()
                                break
                            dec = zlib.decompressobj(wbits=47)
                            at_boundary = False
                            n_members += 1
                        out = dec.decompress(src, READ_BYTES)
                        if dec.eof:
                            src, at_boundary = dec.unused_data, True
                        else:
                            src = dec.unconsumed_tail
                        add("inflate_s", time.perf_counter() - t0)
                        if out:
                            q.put(out)
                q.put(None)
            except BaseException as e:                         # delivered to the consumer
                q.put(e)

        th = threading.Thread(target=producer, daemon=True)
        th.start()
        try:
            while True:
                t0 = time.perf_counter()
                item = q.get()
                add("inflate_wall_s", time.perf_counter() - t0)
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()
            while th.is_alive():                               # unblock a producer waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    th.join(timeout=0.05)


def bgzf_compress(data: bytes, block: int = 0xFF00, level: int = 6) -> bytes:
    """Write `data` as BGZF (what `bgzip` produces): used to make test inputs and by callers who want
    their FASTQ inflatable in parallel."""
    out = []
    for off in list(range(0, len(data), block)) + [None]:
        chunk = b"" if off is None else data[off:off + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        raw = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(raw) + 8
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
                   + raw + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
        if off is None:
            break
    return b"".join(out)
