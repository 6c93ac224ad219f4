"""CPU tests (gloo, world size 2/3) of the batch paths in auriclass_b200/batch.py: query-sharded batch
classification (BASELINE config 4) and 2-D tiled all-vs-all dist (config 5).

The partitioning / gathering is the product code; the per-rank arithmetic is a stand-in engine fed by the
CPU oracle (test infrastructure), because MashGPU needs a GPU.  tests/test_gpu_parity.py runs the same
functions with the real engine on the GPU box.
"""
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


class _Sk:
    def __init__(self, name, h, length):
        self.name, self.h, self.length = name, h, length

    def free(self):
        pass


class _Refs:
    def __init__(self, names, sketches, lengths, k, s):
        self.names, self.sketches, self.lengths, self.k, self.s = names, sketches, lengths, k, s


class OracleEngine:
    """Stand-in for MashGPU: same method names and result layout, arithmetic by the CPU oracle."""

    def sketch_bytes(self, bufs, k=27, s=50000, m=1, mode=1, name="bytes"):
        from oracle import oracle
        sk = oracle.sketch(list(bufs), k, s, m)
        return _Sk(name, sk.hashes, sk.total_len)

    def dist(self, refs, queries):
        from auriclass_b200.mash_gpu import DistRow
        from oracle import oracle
        rows = []
        for q in queries:
            for name, h, length in zip(refs.names, refs.sketches, refs.lengths):
                d, p, cm, dn = oracle.dist(h, q.h, refs.s, refs.k, length, q.length)
                rows.append(DistRow(name, q.name, d, p, cm, dn))
        return rows

    def dist_pairs(self, a, i0, ni, b, j0, nj):
        from oracle import oracle
        D = np.zeros((ni, nj)); P = np.zeros((ni, nj)); CM = np.zeros((ni, nj), np.uint64); DN = np.zeros((ni, nj), np.uint64)
        for i in range(ni):
            for j in range(nj):
                D[i, j], P[i, j], CM[i, j], DN[i, j] = oracle.dist(a.sketches[i0 + i], b.sketches[j0 + j], min(a.s, b.s), a.k,
                                                                    a.lengths[i0 + i], b.lengths[j0 + j])
        return D, P, CM, DN


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _genomes(n, size, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, size, dtype=np.uint8)
    out = []
    for i in range(n):
        g = base.copy()
        mut = rng.random(size) < 0.002 * i
        g[mut] = (g[mut] + rng.integers(1, 4, int(mut.sum()), dtype=np.uint8)) % 4
        seq = np.frombuffer(b"ACGT", np.uint8)[g].tobytes()
        out.append(b">g%d\n" % i + b"\n".join(seq[o:o + 70] for o in range(0, size, 70)) + b"\n")
    return out


def _sketch_set(n, s, seed):
    """C5-style synthetic sketches: clusters sharing a base set, elements replaced at random."""
    rng = np.random.default_rng(seed)
    base = np.unique(rng.integers(0, 1 << 62, 2 * s, dtype=np.uint64))
    out = []
    for i in range(n):
        h = base.copy()
        rep = rng.random(h.size) < rng.uniform(0, 0.3)
        h[rep] = rng.integers(0, 1 << 62, int(rep.sum()), dtype=np.uint64)
        out.append(np.unique(h)[: s - (i % 3)])      # ragged sizes
    return out


def _worker(rank, world, port, fastas, ref_idx, k, s, out_dir):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from auriclass_b200.batch import all_pairs_dense, classify_batch
        eng = OracleEngine()
        ref_sk = [eng.sketch_bytes([fastas[i]], k=k, s=s) for i in ref_idx]
        refs = _Refs([f"ref{i}" for i in ref_idx], [r.h for r in ref_sk], [r.length for r in ref_sk], k, s)
        rows = classify_batch(eng, fastas, refs, k=k, s=s, rank=rank, world=world, dist_batch=2)
        sset = _Refs(None, _sketch_set(11, 300, 7), [10_000 + i for i in range(11)], k, 300)
        dense = all_pairs_dense(eng, sset, 11, tile=4, rank=rank, world=world)
        if rank == 0:
            np.savez(Path(out_dir) / "out.npz", lines=np.array([r.line() for r in rows]), D=dense[0], P=dense[1], CM=dense[2], DN=dense[3])
        else:
            assert rows is None and dense is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_batch_paths_sharded_equal_single_rank(tmp_path, world):
    from auriclass_b200.batch import all_pairs_dense, classify_batch
    k, s = 21, 400
    fastas = _genomes(7, 6_000, seed=world)
    ref_idx = [0, 3, 6]
    mp.spawn(_worker, args=(world, _free_port(), fastas, ref_idx, k, s, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "out.npz")
    eng = OracleEngine()
    ref_sk = [eng.sketch_bytes([fastas[i]], k=k, s=s) for i in ref_idx]
    refs = _Refs([f"ref{i}" for i in ref_idx], [r.h for r in ref_sk], [r.length for r in ref_sk], k, s)
    want = classify_batch(eng, fastas, refs, k=k, s=s)
    assert [r.line() for r in want] == got["lines"].tolist()
    assert len(want) == len(fastas) * len(ref_idx)
    assert [r.query for r in want[:: len(ref_idx)]] == [f"query_{i}" for i in range(len(fastas))]   # query order
    sset = _Refs(None, _sketch_set(11, 300, 7), [10_000 + i for i in range(11)], k, 300)
    D, P, CM, DN = all_pairs_dense(eng, sset, 11, tile=11)    # one tile = brute force
    assert np.array_equal(got["CM"], CM) and np.array_equal(got["DN"], DN)
    assert np.array_equal(got["D"], D) and np.array_equal(got["P"], P)
    assert np.array_equal(CM, CM.T) and np.array_equal(D, D.T)  # Mash's merge is symmetric


def test_tile_plan_covers_upper_triangle_once_and_balances():
    from auriclass_b200.batch import shard_indices, tile_plan
    for n, tile, world in [(0, 4, 2), (1, 4, 8), (10, 3, 4), (50, 7, 8), (16, 16, 3)]:
        plan = tile_plan(n, tile, world)
        cover = np.zeros((n, n), int)
        for t in plan:
            assert 0 <= t.owner < world and t.j0 >= t.i0
            cover[t.i0:t.i0 + t.ni, t.j0:t.j0 + t.nj] += 1
        assert (cover[np.triu_indices(n)] == 1).all()           # every unordered pair (and self pair) exactly once
        per_rank = np.bincount([t.owner for t in plan], minlength=world)
        assert per_rank.max() - per_rank.min() <= 1
        assert sorted(i for r in range(world) for i in shard_indices(n, r, world)) == list(range(n))
    with pytest.raises(ValueError):
        tile_plan(5, 0, 1)
