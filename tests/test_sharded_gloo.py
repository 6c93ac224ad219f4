"""World-size-2/3 CPU tests (gloo) of the multi-GPU exchange protocol in auriclass_b200/sharded.py.

The protocol (tau_cut all-reduce, hash-range all-to-all, owner merge, all-gather, prefix-s) is the
product code; only the per-rank arithmetic is swapped for a numpy stand-in fed by the CPU oracle
(`OracleTableBackend`, test infrastructure), because the CUDA backend needs a GPU.  The merged
sketch must be bit-identical to the oracle's sketch of all reads together.  The same protocol with
the CUDA backend is covered by tests/test_gpu_parity.py::test_sharded_* on the GPU box.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


class _NpSketch:
    def __init__(self, h, c, name):
        self.h, self.c, self.name, self.info = h, c, name, {}


class OracleTableBackend:
    """numpy stand-in for CudaTableBackend: local tables from oracle.all_hashes."""

    device = torch.device("cpu")

    def local_table(self, job, tau_cut_fn, n_ranges):
        from auriclass_b200.sharded import LocalTable, range_bounds
        from oracle import oracle
        k, s, m = job["k"], job["s"], job["m"]
        hs = [oracle.all_hashes(b, k) for b in job["bufs"]]
        allh = np.concatenate(hs) if hs else np.zeros(0, np.uint64)
        if k <= 16:
            allh = allh & np.uint64(0xFFFFFFFF)
        space_max = (1 << 64) - 1 if k > 16 else (1 << 32) - 1
        u, c = np.unique(allh, return_counts=True)
        q = u[c >= m]
        tau = int(q[s - 1]) if q.size >= s else space_max      # what ac_sketch_end_local guarantees
        tau_cut = tau_cut_fn(tau)
        keep = u <= np.uint64(tau_cut)
        u, c = u[keep], c[keep].astype(np.uint32)
        b = np.searchsorted(u, np.array(range_bounds(tau_cut, n_ranges), np.uint64), side="left").tolist() + [u.size]
        splits = [int(b[r + 1] - b[r]) for r in range(n_ranges)]
        counters = (len(job["bufs"]), len(job["bufs"]), 0, int(allh.size), int(keep.sum()))
        return splits, LocalTable(torch.from_numpy(u.view(np.int64).copy()), torch.from_numpy(c.view(np.int32).copy()),
                                  tau_cut, counters)

    def merge_counts(self, hashes, counts, min_copies):
        h = hashes.numpy().view(np.uint64)
        c = counts.numpy().view(np.uint32)
        if h.size == 0:
            return 0
        u, inv = np.unique(h, return_inverse=True)
        tot = np.bincount(inv, weights=c.astype(np.float64)).astype(np.uint32)
        keep = tot >= min_copies
        n = int(keep.sum())
        h[:n] = u[keep]
        c[:n] = tot[keep]
        return n

    def build_sketch(self, hashes, counts, k, s, name):
        return _NpSketch(hashes.numpy().view(np.uint64).copy(), counts.numpy().view(np.uint32).copy(), name)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shards, k, s, m, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from auriclass_b200.sharded import ShardedSketcher
        sh = ShardedSketcher(None, rank, world, backend=OracleTableBackend())
        sk = sh.finish({"bufs": shards[rank], "k": k, "s": s, "m": m}, k, s, m, name="reads")
        np.savez(Path(out_dir) / f"r{rank}.npz", h=sk.h, c=sk.c, n_kmers=sk.info["n_kmers"],
                 tau=np.uint64(sk.info["final_tau"]), sent=sh.last_exchange["sent"], recv=sh.last_exchange["received"])
    finally:
        dist.destroy_process_group()


def _reads(n_genome, n_reads, seed):
    from auriclass_b200 import synth
    rng = np.random.default_rng(seed)
    contigs = [np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n_genome)]]
    return synth.fastq_reads(contigs, n_reads, seed=seed + 1).tobytes()


def _split_records(fq: bytes, world: int):
    lines = fq.split(b"\n")[:-1]
    recs = [b"\n".join(lines[i:i + 4]) + b"\n" for i in range(0, len(lines), 4)]
    per = -(-len(recs) // world)
    return [[b"".join(recs[r * per:(r + 1) * per])] for r in range(world)]


@pytest.mark.parametrize("world,n_genome,n_reads,k,s,m", [
    (2, 20_000, 6_000, 27, 1000, 3),     # truncating: tau_cut < max, >= s qualified
    (2, 3_000, 1_500, 27, 50000, 3),     # never fills: tau_cut = max, all qualified hashes returned
    (3, 20_000, 6_000, 21, 500, 2),      # odd world size, other k / m
    (2, 20_000, 3_000, 16, 1000, 1),     # 32-bit hashes, m = 1
])
def test_sharded_exchange_matches_oracle(tmp_path, world, n_genome, n_reads, k, s, m):
    from oracle import oracle
    fq = _reads(n_genome, n_reads, seed=world * 1000 + k)
    shards = _split_records(fq, world)
    assert b"".join(b for sh in shards for b in sh) == fq
    mp.spawn(_worker, args=(world, _free_port(), shards, k, s, m, str(tmp_path)), nprocs=world, join=True)
    ref = oracle.sketch([fq], k, s, m)
    outs = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for o in outs:
        assert np.array_equal(o["h"], ref.hashes)
        assert np.array_equal(o["c"], ref.counts)
        assert int(o["n_kmers"]) == ref.n_kmers
    assert sum(int(o["sent"]) for o in outs) == sum(int(o["recv"]) for o in outs)


def test_sharded_rank_without_reads(tmp_path):
    """A rank whose shard is empty (fewer read blocks than ranks) still takes part in the collectives."""
    from oracle import oracle
    fq = _reads(10_000, 2_000, seed=5)
    shards = [[fq], [b""]]
    mp.spawn(_worker, args=(2, _free_port(), shards, 27, 800, 2, str(tmp_path)), nprocs=2, join=True)
    ref = oracle.sketch([fq], 27, 800, 2)
    for r in range(2):
        o = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(o["h"], ref.hashes) and np.array_equal(o["c"], ref.counts)


def test_ordered_i64_mapping_and_bounds():
    from auriclass_b200.sharded import ordered_i64_to_u64, range_bounds, u64_to_ordered_i64
    xs = [0, 1, (1 << 63) - 1, 1 << 63, (1 << 64) - 1, 0x0123456789ABCDEF]
    ys = [u64_to_ordered_i64(x) for x in xs]
    assert [ordered_i64_to_u64(y) for y in ys] == xs
    assert sorted(ys) == [u64_to_ordered_i64(x) for x in sorted(xs)]
    for tau in (0, 7, (1 << 64) - 1, (1 << 60) + 12345):
        for P in (1, 2, 3, 8):
            b = range_bounds(tau, P)
            assert b[0] == 0 and all(b[i] <= b[i + 1] for i in range(P - 1))
            for x in {0, tau, tau // 2, tau // 3}:   # owner(x) = floor(x P / (tau+1)) lands in its own range
                r = x * P // (tau + 1)
                assert b[r] <= x and (r == P - 1 or x < b[r + 1])
