"""Synthetic C. auris-sized inputs: a random ~12.4 Mb genome in 7
contigs, 150 bp reads with substitution errors and N, 4-line FASTQ / 80-column FASTA text.
Used by bench.py and the size-independent parity tests; numpy only (deterministic PCG64)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

CONTIGS_MB = (4.2, 2.3, 1.6, 1.4, 1.2, 1.0, 0.7)  # 12.4 Mb
_ACGT = np.frombuffer(b"ACGT", np.uint8)
_COMP = np.zeros(256, np.uint8)
_COMP[list(b"ACGTN")] = list(b"TGCAN")


def genome(seed: int = 20240001, scale: float = 1.0) -> List[np.ndarray]:
    """i.i.d. uniform ACGT contigs; `scale` shrinks every contig (tests)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return [_ACGT[rng.integers(0, 4, int(mb * 1e6 * scale), dtype=np.uint8)] for mb in CONTIGS_MB]


def mutate(contigs: List[np.ndarray], rate: float, seed: int) -> List[np.ndarray]:
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for c in contigs:
        c = c.copy()
        n = rng.binomial(c.size, rate)
        pos = rng.integers(0, c.size, n)
        c[pos] = _ACGT[rng.integers(0, 4, n, dtype=np.uint8)]
        out.append(c)
    return out


def fasta_bytes(contigs: List[np.ndarray], width: int = 80, prefix: str = "synth_caur") -> bytes:
    parts = []
    for i, c in enumerate(contigs):
        parts.append(f">{prefix}_{i}\n".encode())
        full = (c.size // width) * width
        body = np.empty((c.size // width, width + 1), np.uint8)
        body[:, :width] = c[:full].reshape(-1, width)
        body[:, width] = 10
        parts.append(body.tobytes())
        if full < c.size:
            parts.append(c[full:].tobytes() + b"\n")
    return b"".join(parts)


def fastq_reads(contigs: List[np.ndarray], n_reads: int, seed: int, read_len: int = 150, sub_rate: float = 0.005,
                n_rate: float = 0.0002, file_tag: int = 0, out: np.ndarray = None, block: int = 1 << 18) -> np.ndarray:
    """4-line FASTQ text as a uint8 array: `@SYN{tag}.{idx:09d}/1`, sequence, `+`, quality 'I'*len.
    Reads start uniformly over the concatenated genome (never spanning contigs), random strand."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.array([c.size for c in contigs])
    usable = np.maximum(lens - read_len + 1, 0)
    cum = np.concatenate([[0], np.cumsum(usable)])
    cat = np.concatenate(contigs)
    coff = np.concatenate([[0], np.cumsum(lens)])[:-1]
    hdr_len = len(f"@SYN{file_tag}.{0:09d}/1")
    rec = hdr_len + 1 + read_len + 3 + read_len + 1
    if out is None:
        out = np.empty(n_reads * rec, np.uint8)
    assert out.size >= n_reads * rec
    view = out[: n_reads * rec].reshape(n_reads, rec)
    ar = np.arange(read_len)
    digits = 10 ** np.arange(8, -1, -1)
    hdr_prefix = np.frombuffer(f"@SYN{file_tag}.".encode(), np.uint8)
    for b0 in range(0, n_reads, block):
        b1 = min(n_reads, b0 + block)
        nb = b1 - b0
        u = rng.integers(0, cum[-1], nb)
        ci = np.searchsorted(cum, u, side="right") - 1
        start = coff[ci] + (u - cum[ci])
        seq = cat[start[:, None] + ar[None, :]]
        rev = rng.random(nb) < 0.5
        seq[rev] = _COMP[seq[rev][:, ::-1]]
        sub = rng.random((nb, read_len)) < sub_rate
        nsub = int(sub.sum())
        seq[sub] = _ACGT[rng.integers(0, 4, nsub, dtype=np.uint8)]
        nn = rng.random((nb, read_len)) < n_rate
        seq[nn] = ord("N")
        v = view[b0:b1]
        v[:, : hdr_prefix.size] = hdr_prefix
        idx = np.arange(b0, b1)
        v[:, hdr_prefix.size : hdr_prefix.size + 9] = (idx[:, None] // digits[None, :]) % 10 + 48
        v[:, hdr_prefix.size + 9 : hdr_len] = np.frombuffer(b"/1", np.uint8)
        p = hdr_len
        v[:, p] = 10
        v[:, p + 1 : p + 1 + read_len] = seq
        p += 1 + read_len
        v[:, p : p + 3] = np.frombuffer(b"\n+\n", np.uint8)
        v[:, p + 3 : p + 3 + read_len] = ord("I")
        v[:, p + 3 + read_len] = 10
    return out[: n_reads * rec]


def record_bytes(read_len: int = 150, file_tag: int = 0) -> int:
    return len(f"@SYN{file_tag}.{0:09d}/1") + 1 + read_len + 3 + read_len + 1
