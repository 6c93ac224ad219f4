"""Reader / writer for Mash `.msh` sketch files (unpacked Cap'n Proto stream).

Host-side duty (1) of BASELINE.json's north_star: the reference sketch that
`BasicAuriclass.run_mash_dist` hands to `mash dist`
(/root/reference/auriclass/classes.py:92-97, `reference_sketch_path`) is parsed
here and uploaded once through `ac_refs_load`.  Layout follows SURVEY.md App. B,
derived by walking /root/reference/tests/data/ref_sketch.msh; pycapnp is not
available (and not needed): the subset of the encoding Mash emits is small.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

_DEFAULT_SEED = 42  # MinHash.capnp: hashSeed :UInt32 = 42 (stored XOR default)


@dataclass
class MshReference:
    name: str
    comment: str
    length: int
    hashes: np.ndarray  # uint64, ascending (32-bit sketches are zero-extended)
    counts: Optional[np.ndarray] = None  # uint32 or None


@dataclass
class MshSketch:
    kmer_size: int
    sketch_size: int
    seed: int = _DEFAULT_SEED
    hash_bits: int = 64
    alphabet: str = "ACGT"
    concatenated: bool = True
    noncanonical: bool = False
    preserve_case: bool = False
    error: float = 0.0
    window_size: int = 0
    references: List[MshReference] = field(default_factory=list)


class MshFormatError(ValueError):
    pass


class _Msg:
    """Minimal Cap'n Proto message walker (struct / list / far pointers)."""

    def __init__(self, raw: bytes):
        if len(raw) < 8:
            raise MshFormatError("file too short for a Cap'n Proto stream header")
        nseg = struct.unpack_from("<I", raw, 0)[0] + 1
        if nseg > 1 << 16:
            raise MshFormatError("implausible segment count (packed or not a .msh?)")
        sizes = struct.unpack_from(f"<{nseg}I", raw, 4)
        off = 4 + 4 * nseg
        off += (-off) % 8
        self.segs: List[memoryview] = []
        mv = memoryview(raw)
        for sz in sizes:
            end = off + 8 * sz
            if end > len(raw):
                raise MshFormatError("segment table exceeds file size")
            self.segs.append(mv[off:end])
            off = end

    def word(self, seg: int, w: int) -> int:
        return struct.unpack_from("<Q", self.segs[seg], 8 * w)[0]

    def resolve(self, seg: int, w: int) -> Tuple[int, int, int]:
        """Follow far pointers. Returns (ptr_word, seg, target_base_word) where the
        content starts at target_base_word + 1 + signed offset of ptr_word."""
        p = self.word(seg, w)
        if p & 3 != 2:
            return p, seg, w
        double = (p >> 2) & 1
        pad_off = (p >> 3) & 0x1FFFFFFF
        pad_seg = p >> 32
        if not double:
            return self.word(pad_seg, pad_off), pad_seg, pad_off
        # double-far: pad = [far ptr to content start, tag word with offset 0]
        far2 = self.word(pad_seg, pad_off)
        tag = self.word(pad_seg, pad_off + 1)
        cseg = far2 >> 32
        cstart = (far2 >> 3) & 0x1FFFFFFF
        # content begins exactly at cstart: emulate "base + 1 + 0"
        return tag & ~0xFFFFFFFC, cseg, cstart - 1

    @staticmethod
    def _soff(p: int) -> int:
        o = (p >> 2) & 0x3FFFFFFF
        return o - (1 << 30) if o & (1 << 29) else o

    def struct_at(self, seg: int, w: int):
        p, s, base = self.resolve(seg, w)
        if p == 0:
            return None
        if p & 3 != 0:
            raise MshFormatError("expected struct pointer")
        start = base + 1 + self._soff(p)
        return (s, start, (p >> 32) & 0xFFFF, (p >> 48) & 0xFFFF)

    def list_at(self, seg: int, w: int):
        p, s, base = self.resolve(seg, w)
        if p == 0:
            return None
        if p & 3 != 1:
            raise MshFormatError("expected list pointer")
        start = base + 1 + self._soff(p)
        return (s, start, (p >> 32) & 7, p >> 35)

    def text_at(self, seg: int, w: int) -> str:
        l = self.list_at(seg, w)
        if l is None:
            return ""
        s, start, code, n = l
        if code != 2:
            raise MshFormatError("Text must be a byte list")
        b = bytes(self.segs[s][8 * start : 8 * start + n])
        return b[:-1].decode("utf-8") if n else ""

    def prim_list(self, seg: int, w: int, dtype) -> Optional[np.ndarray]:
        l = self.list_at(seg, w)
        if l is None:
            return None
        s, start, code, n = l
        want = {4: 4, 8: 5}[np.dtype(dtype).itemsize]
        if code != want:
            raise MshFormatError(f"list element size code {code}, expected {want}")
        return np.frombuffer(self.segs[s], dtype=dtype, count=n, offset=8 * start).copy()


def read_msh(path: Union[str, Path]) -> MshSketch:
    raw = Path(path).read_bytes()
    m = _Msg(raw)
    root = m.struct_at(0, 0)
    if root is None:
        raise MshFormatError("null root pointer")
    seg, ds, ndata, nptr = root
    if ndata < 2 or nptr < 3:
        raise MshFormatError("root struct too small for MinHash")
    w0, w1 = m.word(seg, ds), m.word(seg, ds + 1)
    w2 = m.word(seg, ds + 2) if ndata > 2 else 0
    out = MshSketch(
        kmer_size=w0 & 0xFFFFFFFF,
        window_size=w0 >> 32,
        sketch_size=w1 & 0xFFFFFFFF,
        concatenated=bool((w1 >> 32) & 1),
        noncanonical=bool((w1 >> 33) & 1),
        preserve_case=bool((w1 >> 34) & 1),
        error=struct.unpack("<f", struct.pack("<I", w2 & 0xFFFFFFFF))[0],
        seed=((w2 >> 32) & 0xFFFFFFFF) ^ _DEFAULT_SEED,
    )
    pbase = ds + ndata
    out.alphabet = m.text_at(seg, pbase + 2) or "ACGT"
    # referenceListOld (ptr0) is used when seed == 42, referenceList (ptr3) otherwise;
    # take whichever is populated.
    refs = None
    for pi in (0, 3):
        if pi >= nptr:
            continue
        rl = m.struct_at(seg, pbase + pi)
        if rl is None:
            continue
        rseg, rds, rnd, rnp = rl
        if rnp < 1:
            continue
        lst = m.list_at(rseg, rds + rnd)
        if lst is None:
            continue
        lseg, lstart, code, nwords = lst
        if code != 7:
            raise MshFormatError("reference list must be an inline-composite list")
        tag = m.word(lseg, lstart)
        count = (tag >> 2) & 0x3FFFFFFF
        edata, eptr = (tag >> 32) & 0xFFFF, (tag >> 48) & 0xFFFF
        if count == 0:
            continue
        refs = (lseg, lstart + 1, count, edata, eptr)
        break
    use64 = None
    if refs is not None:
        lseg, first, count, edata, eptr = refs
        for i in range(count):
            es = first + i * (edata + eptr)
            length32 = m.word(lseg, es) & 0xFFFFFFFF if edata > 0 else 0
            length64 = m.word(lseg, es + 1) if edata > 1 else 0
            pb = es + edata
            name = m.text_at(lseg, pb + 2) if eptr > 2 else ""
            comment = m.text_at(lseg, pb + 3) if eptr > 3 else ""
            h32 = m.prim_list(lseg, pb + 4, np.uint32) if eptr > 4 else None
            h64 = m.prim_list(lseg, pb + 5, np.uint64) if eptr > 5 else None
            c32 = m.prim_list(lseg, pb + 6, np.uint32) if eptr > 6 else None
            if h64 is not None and len(h64):
                hashes, this64 = h64, True
            elif h32 is not None and len(h32):
                hashes, this64 = h32.astype(np.uint64), False
            else:
                hashes, this64 = np.zeros(0, np.uint64), out.kmer_size > 16
            use64 = this64 if use64 is None else use64
            out.references.append(
                MshReference(name, comment, int(length64 or length32), hashes,
                             c32 if c32 is not None and len(c32) else None)
            )
    out.hash_bits = 64 if (use64 if use64 is not None else out.kmer_size > 16) else 32
    return out


# --------------------------------------------------------------------------
# Writer: one segment, no far pointers (valid input for `mash info/dist`).
# --------------------------------------------------------------------------
class _Builder:
    def __init__(self):
        self.words = bytearray()

    def alloc(self, n: int) -> int:
        w = len(self.words) // 8
        self.words.extend(b"\0" * (8 * n))
        return w

    def put(self, w: int, v: int) -> None:
        struct.pack_into("<Q", self.words, 8 * w, v & 0xFFFFFFFFFFFFFFFF)

    def put_bytes(self, w: int, b: bytes) -> None:
        self.words[8 * w : 8 * w + len(b)] = b

    def struct_ptr(self, at: int, target: int, nd: int, npt: int) -> None:
        off = target - at - 1
        self.put(at, ((off & 0x3FFFFFFF) << 2) | (nd << 32) | (npt << 48))

    def list_ptr(self, at: int, target: int, code: int, n: int) -> None:
        off = target - at - 1
        self.put(at, 1 | ((off & 0x3FFFFFFF) << 2) | (code << 32) | (n << 35))

    def text(self, at: int, s: str) -> None:
        b = s.encode("utf-8") + b"\0"
        t = self.alloc((len(b) + 7) // 8)
        self.put_bytes(t, b)
        self.list_ptr(at, t, 2, len(b))

    def prim(self, at: int, arr: np.ndarray, code: int) -> None:
        b = arr.tobytes()
        t = self.alloc((len(b) + 7) // 8)
        self.put_bytes(t, b)
        self.list_ptr(at, t, code, len(arr))


def write_msh(path: Union[str, Path], sk: MshSketch) -> None:
    b = _Builder()
    root = b.alloc(1)
    ms = b.alloc(3 + 4)
    b.struct_ptr(root, ms, 3, 4)
    b.put(ms, (sk.kmer_size & 0xFFFFFFFF) | (sk.window_size << 32))
    flags = int(sk.concatenated) | (int(sk.noncanonical) << 1) | (int(sk.preserve_case) << 2)
    b.put(ms + 1, (sk.sketch_size & 0xFFFFFFFF) | (flags << 32))
    err_bits = struct.unpack("<I", struct.pack("<f", sk.error))[0]
    b.put(ms + 2, err_bits | (((sk.seed ^ _DEFAULT_SEED) & 0xFFFFFFFF) << 32))
    p = ms + 3
    # reference list holder struct {references @0}
    holder = b.alloc(1)
    b.struct_ptr(p + (0 if sk.seed == _DEFAULT_SEED else 3), holder, 0, 1)
    n = len(sk.references)
    ED, EP = 3, 7
    tagw = b.alloc(1 + n * (ED + EP))
    b.put(tagw, (n << 2) | (ED << 32) | (EP << 48))
    b.list_ptr(holder, tagw, 7, n * (ED + EP))
    for i, r in enumerate(sk.references):
        es = tagw + 1 + i * (ED + EP)
        b.put(es, r.length if r.length < (1 << 32) else 0)
        b.put(es + 1, r.length)
        b.put(es + 2, 0)
        pb = es + ED
        b.text(pb + 2, r.name)
        b.text(pb + 3, r.comment)
        if sk.hash_bits == 64:
            b.prim(pb + 5, np.ascontiguousarray(r.hashes, dtype=np.uint64), 5)
        else:
            b.prim(pb + 4, np.ascontiguousarray(r.hashes, dtype=np.uint64).astype(np.uint32), 4)
        if r.counts is not None:
            b.prim(pb + 6, np.ascontiguousarray(r.counts, dtype=np.uint32), 4)
    # empty locus list (composite, zero elements) keeps `mash info` happy
    lt = b.alloc(1)
    b.put(lt, (0 << 2) | (1 << 32) | (0 << 48))
    b.list_ptr(p + 1, lt, 7, 0)
    b.text(p + 2, sk.alphabet)
    nwords = len(b.words) // 8
    with open(path, "wb") as f:
        f.write(struct.pack("<II", 0, nwords))
        f.write(bytes(b.words))


def pack_references(refs: Sequence[MshReference]):
    """Concatenate reference hash lists for `ac_refs_load`: (hashes, offsets, lengths)."""
    offsets = np.zeros(len(refs) + 1, dtype=np.uint64)
    for i, r in enumerate(refs):
        offsets[i + 1] = offsets[i] + len(r.hashes)
    hashes = (np.concatenate([np.ascontiguousarray(r.hashes, dtype=np.uint64) for r in refs])
              if refs else np.zeros(0, np.uint64))
    lengths = np.array([r.length for r in refs], dtype=np.uint64)
    return hashes, offsets, lengths
