"""Read blocks: the unit by which `mash sketch -r -m` input is sharded over the GPUs of one box.

SURVEY.md section 8e / BASELINE.json configs[2]: "input bytes cut at record boundaries into read blocks,
block b -> GPU b mod P".  The reference hands all read files to ONE `mash sketch` process
(/root/reference/auriclass/classes.py:576-589, `*self.read_paths`); the sketch is a set function of the reads
(SURVEY.md section 0 fact 9), so any partition of the records over ranks gives the same result once the counts
are merged (sharded.py).

Every rank runs the same deterministic cutter over the same inflated byte stream and keeps the blocks whose
global index is congruent to its rank; nothing is exchanged to agree on the cut.  A cut is placed at the first
*record start* at or after every multiple of `block_bytes` of a file.  For strict 4-line FASTQ (the only FASTQ the
GPU parser accepts) a record start can be recognised locally, without knowing the line phase of the position:

    a line that starts with '@' and whose next-but-one line starts with '+'

is a header line -- the only other line that may start with '@' is a quality line, and the line two below a quality
line is a sequence line, which cannot start with '+'.  So the cutter never has to count newlines from the start of
the file, and a rank that can seek (plain files) could jump straight to its blocks.
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

DEFAULT_BLOCK_BYTES = 32 << 20


def find_record_start(buf, start: int, line_start0: bool = True) -> int:
    """Smallest p >= start at which a FASTQ record starts in `buf` (bytes-like), or -1 when `buf` ends before one can
    be confirmed (the first character of the line after next must be inside `buf`).  `line_start0`: buf[0] begins a
    line (the byte before it is a newline or the start of the file)."""
    mv = bytes(buf) if not isinstance(buf, (bytes, bytearray)) else buf
    n = len(mv)
    p = start
    while p < n:
        at_line_start = line_start0 if p == 0 else mv[p - 1] == 0x0A
        if not at_line_start or mv[p] != 0x40:
            q = mv.find(b"\n@", p - 1 if p > 0 and not at_line_start else p)
            if q < 0:
                return -1
            p = q + 1
        # p is at a line start and holds '@': look at the line after next
        e1 = mv.find(b"\n", p)
        if e1 < 0:
            return -1
        e2 = mv.find(b"\n", e1 + 1)
        if e2 < 0 or e2 + 1 >= n:
            return -1
        if mv[e2 + 1] == 0x2B:
            return p
        p += 1      # a quality line that starts with '@': keep looking
    return -1


class ReadBlockCutter:
    """Streaming cutter: feed the inflated chunks of the files in order, get (block index, piece) pairs back.
    A block may arrive as several pieces; pieces of one block are consecutive.  Block indices run over all files."""

    def __init__(self, block_bytes: int = DEFAULT_BLOCK_BYTES):
        if block_bytes < 1:
            raise ValueError("block_bytes must be positive")
        self.block_bytes = int(block_bytes)
        self.block_idx = -1          # index of the block being filled
        self._filled = 0             # bytes of the current block handed out so far
        self._held = b""             # bytes past the nominal block size, waiting for a confirmed record start
        self._line_start = True      # the next byte to be looked at begins a line

    def start_file(self) -> None:
        """A file always starts a new block (k-mers never span files: SURVEY.md section 8b)."""
        if self._held:
            raise RuntimeError("start_file() before end_file()")
        self.block_idx += 1
        self._filled = 0
        self._line_start = True

    def _emit(self, piece: bytes) -> Tuple[int, bytes]:
        self._filled += len(piece)
        self._line_start = piece[-1] == 0x0A
        return self.block_idx, piece

    def _cut(self, data: bytes, last: bool) -> Iterator[Tuple[int, bytes]]:
        off = 0
        while off < len(data):
            room = self.block_bytes - self._filled
            if room > 0:
                take = min(room, len(data) - off)
                yield self._emit(data[off:off + take])
                off += take
                continue
            # the block is full: it ends at the next record start
            p = find_record_start(data[off:] if off else data, 0, self._line_start)
            if p < 0:
                if last:
                    yield self._emit(data[off:])
                    return
                # Not decidable yet.  A candidate needs its two following newlines and one more byte inside the buffer,
                # so every line start up to the third-last newline has been looked at and rejected: those bytes belong
                # to this block; only the tail behind that newline stays undecided.
                keep = len(data)
                for _ in range(3):
                    keep = data.rfind(b"\n", off, keep)
                    if keep < 0:
                        break
                if keep >= off:
                    yield self._emit(data[off:keep + 1])
                    off = keep + 1
                self._held = data[off:]
                return
            if p > 0:
                yield self._emit(data[off:off + p])
            off += p
            self.block_idx += 1
            self._filled = 0
            self._line_start = True

    def feed(self, chunk) -> Iterator[Tuple[int, bytes]]:
        data = self._held + bytes(chunk) if self._held else bytes(chunk)
        self._held = b""
        yield from self._cut(data, False)

    def end_file(self) -> Iterator[Tuple[int, bytes]]:
        data, self._held = self._held, b""
        if data:
            yield from self._cut(data, True)


def cut_read_blocks(files: Sequence, block_bytes: int = DEFAULT_BLOCK_BYTES) -> List[Tuple[int, int, int]]:
    """In-memory form for whole buffers (bench.py, tests): [(file index, begin, end)] of every read block, in
    block-index order.  Same cuts as the streaming cutter."""
    out: List[Tuple[int, int, int]] = []
    for fi, buf in enumerate(files):
        mv = buf if isinstance(buf, (bytes, bytearray)) else np.ascontiguousarray(buf, np.uint8)
        n = len(mv)
        begin = 0
        if n == 0:
            out.append((fi, 0, 0))
            continue
        while begin < n:
            want = begin + block_bytes
            if want >= n:
                out.append((fi, begin, n))
                break
            # a window that certainly holds the next record start (grown if a line is longer than expected)
            win = 1 << 16
            while True:
                hi = min(n, want + win)
                seg = bytes(memoryview(mv)[want - 1:hi]) if want > 0 else bytes(memoryview(mv)[:hi])
                p = find_record_start(seg, 1 if want > 0 else 0)
                if p >= 0 or hi == n:
                    break
                win *= 4
            if p < 0:
                out.append((fi, begin, n))
                break
            end = want - 1 + p if want > 0 else p
            out.append((fi, begin, end))
            begin = end
    return out


def blocks_of_rank(blocks: Iterable[Tuple[int, int, int]], rank: int, world: int) -> List[Tuple[int, int, int]]:
    """Block b -> rank b mod P."""
    return [b for i, b in enumerate(blocks) if i % world == rank]
