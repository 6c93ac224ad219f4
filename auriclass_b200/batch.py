"""The two batch paths of SURVEY.md section 8(e) that shard by independent units (no data-path collective):

* **batch classification** (BASELINE.json configs[3]): many assemblies, each sketched (`mash sketch`, one
  sketch per file: /root/reference/auriclass/classes.py:663-723) and distanced against the clade references
  (`mash dist REF QUERY`: classes.py:67-119).  Query q belongs to rank q mod P; the reference sketches are
  replicated (9.6 MB).  The 24 rows per query are gathered on rank 0 by the host.
* **all-vs-all `mash dist`** (configs[4]): the sketch set is replicated on every GPU (50,000 x 400 KB = 20 GB
  of 180 GB), the upper triangle of the pair matrix is cut into 2-D tiles dealt round-robin over the ranks
  (`ac_dist_pairs`, one warp per pair), results stay sharded per rank.  Mash's merge is symmetric in its two
  operands, so the lower triangle is the mirror image.

Only the host-side partitioning lives here; the arithmetic is `MashGPU.sketch_* / dist / dist_pairs`
(libauriclass_cuda.so).  The same logic runs in tests/test_batch_gloo.py with a stand-in engine.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np

from .mash_gpu import DistRow

PathOrBytes = Union[str, bytes, "os.PathLike[str]"]


# ------------------------------------------------------------------ C4: batch classification
def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """Query q -> rank q mod P."""
    return list(range(rank, n, world))


def classify_batch(engine, queries: Sequence[PathOrBytes], refs, k: int = 27, s: int = 50000, *, rank: int = 0,
                   world: int = 1, group=None, names: Optional[Sequence[str]] = None,
                   dist_batch: int = 64) -> Optional[List[DistRow]]:
    """Sketch this rank's share of `queries` (paths, or in-memory FASTA bytes) and distance each against
    `refs`.  Returns, on rank 0, the rows of all queries in query order (24 per query, references in file
    order -- what `mash dist REF q0 q1 ...` prints); None on the other ranks."""
    from ._lib import AC_MODE_FASTA
    mine = shard_indices(len(queries), rank, world)
    rows_by_q: Dict[int, List[DistRow]] = {}
    pending = []

    def flush():
        if not pending:
            return
        rows = engine.dist(refs, [sk for _, sk in pending])
        nr = len(refs.names)
        for n, (qi, sk) in enumerate(pending):
            rows_by_q[qi] = rows[n * nr:(n + 1) * nr]
            sk.free()
        pending.clear()

    for qi in mine:
        q = queries[qi]
        name = names[qi] if names else (f"query_{qi}" if isinstance(q, (bytes, bytearray, memoryview)) else str(q))
        if isinstance(q, (bytes, bytearray, memoryview)):
            sk = engine.sketch_bytes([bytes(q)], k=k, s=s, m=1, mode=AC_MODE_FASTA, name=name)
        else:
            sk = engine.sketch_fasta(q, k=k, s=s)
            sk.name = name
        pending.append((qi, sk))
        if len(pending) >= dist_batch:
            flush()
    flush()
    if world == 1:
        return [r for qi in sorted(rows_by_q) for r in rows_by_q[qi]]
    import torch.distributed as dist
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(rows_by_q, gathered, dst=0, group=group)
    if rank != 0:
        return None
    merged: Dict[int, List[DistRow]] = {}
    for part in gathered:
        merged.update(part)
    assert len(merged) == len(queries), "every query must come back exactly once"
    return [r for qi in range(len(queries)) for r in merged[qi]]


# ------------------------------------------------------------------ C5: all-vs-all dist
@dataclass(frozen=True)
class Tile:
    i0: int
    ni: int
    j0: int
    nj: int
    owner: int


def tile_plan(n: int, tile: int, world: int) -> List[Tile]:
    """Upper triangle (diagonal tiles included) of the n x n pair matrix in tile x tile blocks, dealt
    round-robin over the ranks in row-major tile order."""
    if n < 0 or tile < 1 or world < 1:
        raise ValueError("tile_plan: n >= 0, tile >= 1, world >= 1")
    out, t = [], 0
    for i0 in range(0, n, tile):
        for j0 in range(i0, n, tile):
            out.append(Tile(i0, min(tile, n - i0), j0, min(tile, n - j0), t % world))
            t += 1
    return out


@dataclass
class TileResult:
    tile: Tile
    distance: np.ndarray   # [ni, nj] float64
    pvalue: np.ndarray     # [ni, nj] float64
    common: np.ndarray     # [ni, nj] uint64
    denom: np.ndarray      # [ni, nj] uint64


def all_pairs(engine, sketch_set, n: int, *, tile: int = 512, rank: int = 0, world: int = 1) -> Iterator[TileResult]:
    """This rank's tiles of `mash dist SET SET` (results stay sharded: the caller consumes them tile by tile).
    Entries below the diagonal inside a diagonal tile are computed too (they cost nothing extra to keep the
    tile rectangular) and equal their mirror images."""
    for t in tile_plan(n, tile, world):
        if t.owner != rank:
            continue
        d, p, cm, dn = engine.dist_pairs(sketch_set, t.i0, t.ni, sketch_set, t.j0, t.nj)
        yield TileResult(t, d, p, cm, dn)


def all_pairs_dense(engine, sketch_set, n: int, *, tile: int = 512, rank: int = 0, world: int = 1, group=None):
    """Small-n convenience (tests, reports): the full symmetric matrices on rank 0, None elsewhere."""
    mine = list(all_pairs(engine, sketch_set, n, tile=tile, rank=rank, world=world))
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0, group=group)
        if rank != 0:
            return None
        mine = [r for part in gathered for r in part]
    D = np.zeros((n, n), np.float64); P = np.zeros((n, n), np.float64)
    CM = np.zeros((n, n), np.uint64); DN = np.zeros((n, n), np.uint64)
    seen = np.zeros((n, n), bool)
    for r in mine:
        t = r.tile
        sl = (slice(t.i0, t.i0 + t.ni), slice(t.j0, t.j0 + t.nj))
        D[sl], P[sl], CM[sl], DN[sl] = r.distance, r.pvalue, r.common, r.denom
        seen[sl] = True
        slT = (sl[1], sl[0])
        if t.i0 != t.j0:
            D[slT], P[slT], CM[slT], DN[slT] = r.distance.T, r.pvalue.T, r.common.T, r.denom.T
            seen[slT] = True
    assert seen.all(), "tile plan must cover the whole matrix"
    return D, P, CM, DN
