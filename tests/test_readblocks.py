"""Read-block cutting (auriclass_b200/readblocks.py; SURVEY.md section 8e: "input bytes cut at record boundaries into
read blocks, block b -> GPU b mod P") and its use by the sharded FASTQ path, without a GPU:

* hypothesis: for arbitrary 4-line FASTQ (ragged reads, '@' / '+' at the start of quality lines and inside headers,
  missing final newline, empty files) every cut is a true record start, the blocks tile each file, and the streaming
  cutter fed with arbitrary chunkings makes exactly the cuts of the in-memory one;
* gloo, world 2 and 3: ranks that each keep the blocks b = rank mod P of two read files and merge through
  sharded.ShardedSketcher reproduce the oracle's sketch of all reads (the reference passes all files to one
  `mash sketch -r`, /root/reference/auriclass/classes.py:576-589);
* a rank that fails before the exchange makes every rank raise instead of hanging in the first collective.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp
from hypothesis import given, settings
from hypothesis import strategies as st

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from auriclass_b200.readblocks import ReadBlockCutter, blocks_of_rank, cut_read_blocks, find_record_start  # noqa: E402
from tests.test_sharded_gloo import OracleTableBackend, _free_port, _reads  # noqa: E402

seq_chars = st.text(alphabet="ACGTNacgtn", min_size=1, max_size=60)
qual_first = st.sampled_from("@+I#5!~")


@st.composite
def fastq_record(draw):
    seq = draw(seq_chars)
    hdr = "@" + draw(st.text(alphabet="rd@+:/ 0123456789", max_size=12))
    plus = "+" + draw(st.sampled_from(["", "", "again", "@x"]))
    qual = draw(qual_first) + draw(st.text(alphabet="@+I#5!~", min_size=len(seq) - 1, max_size=len(seq) - 1)) if len(seq) else ""
    return f"{hdr}\n{seq}\n{plus}\n{qual}\n".encode()


fastq_file = st.lists(fastq_record(), max_size=25).map(b"".join)


@settings(max_examples=300, deadline=None)
@given(files=st.lists(fastq_file, min_size=1, max_size=3), block=st.integers(1, 400), strip_nl=st.booleans(),
       chunks=st.lists(st.integers(1, 97), min_size=1, max_size=8))
def test_cuts_are_record_starts_and_streaming_equals_in_memory(files, block, strip_nl, chunks):
    if strip_nl and files[-1]:
        files = files[:-1] + [files[-1][:-1]]
    blocks = cut_read_blocks(files, block)
    for fi, f in enumerate(files):
        mine = [b for b in blocks if b[0] == fi]
        assert mine[0][1] == 0 and mine[-1][2] == len(f)
        assert all(x[2] == y[1] for x, y in zip(mine, mine[1:]))
        for (_, a0, _), (_, a, _) in zip(mine, mine[1:]):
            assert f[:a].count(b"\n") % 4 == 0 and f[a - 1:a] == b"\n" and f[a:a + 1] == b"@"
            assert a - a0 >= block                       # a block is cut at the FIRST record start past its size
            assert find_record_start(f, a0 + block) == a
    cut = ReadBlockCutter(block)
    got = {}
    ci = 0
    for f in files:
        cut.start_file()
        off = 0
        while off < len(f):
            n = chunks[ci % len(chunks)]
            ci += 1
            for bi, piece in cut.feed(f[off:off + n]):
                got.setdefault(bi, []).append(piece)
            off += n
        for bi, piece in cut.end_file():
            got.setdefault(bi, []).append(piece)
    assert [b"".join(got.get(i, [])) for i in range(len(blocks))] == [bytes(files[fi][a:b]) for fi, a, b in blocks]
    # dealing the blocks over P ranks is a partition
    for world in (1, 2, 3):
        parts = [blocks_of_rank(blocks, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(blocks)


def _worker(rank, world, port, files, block, k, s, m, out_dir, fail_rank):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from auriclass_b200 import _lib
        from auriclass_b200.sharded import PeerFailure, ShardedSketcher
        sh = ShardedSketcher(None, rank, world, backend=OracleTableBackend())
        # what MashGPU.sketch_reads_sharded does on the host: one cutter over all files, keep blocks b = rank mod P
        cut = ReadBlockCutter(block)
        mine = []
        for f in files:
            cut.start_file()
            cur = []
            for off in range(0, len(f), 10_007):
                cur += [p for bi, p in cut.feed(f[off:off + 10_007]) if bi % world == rank]
            cur += [p for bi, p in cut.end_file() if bi % world == rank]
            mine.append(b"".join(cur))
        if fail_rank is not None:
            try:
                sh.finish(0, k, s, m, failure=ValueError("bad block") if rank == fail_rank else None) if rank == fail_rank \
                    else sh.finish({"bufs": mine, "k": k, "s": s, "m": m}, k, s, m)
                outcome = "returned"
            except ValueError:
                outcome = "own"
            except PeerFailure as e:
                outcome = "peer:" + str(e.msg)
            (Path(out_dir) / f"r{rank}.txt").write_text(outcome)
            return
        sk = sh.finish({"bufs": mine, "k": k, "s": s, "m": m}, k, s, m, name="reads")
        np.savez(Path(out_dir) / f"r{rank}.npz", h=sk.h, c=sk.c, nbytes=sum(len(x) for x in mine))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,block", [(2, 40_000), (3, 25_000), (2, 10_000_000)])
def test_read_block_sharding_over_gloo_matches_oracle(tmp_path, world, block):
    from oracle import oracle
    files = [_reads(15_000, 2_500, seed=11), _reads(15_000, 2_000, seed=11)[:-1]]   # R1, R2 (no final newline)
    mp.spawn(_worker, args=(world, _free_port(), files, block, 27, 700, 3, str(tmp_path), None), nprocs=world, join=True)
    ref = oracle.sketch(files, 27, 700, 3)
    outs = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    assert sum(int(o["nbytes"]) for o in outs) == sum(len(f) for f in files)
    if block < 100_000:
        assert all(int(o["nbytes"]) > 0 for o in outs)          # every rank really had blocks
    for o in outs:
        assert np.array_equal(o["h"], ref.hashes) and np.array_equal(o["c"], ref.counts)


def test_a_failing_rank_makes_every_rank_raise(tmp_path):
    files = [_reads(5_000, 500, seed=3)]
    mp.spawn(_worker, args=(3, _free_port(), files, 20_000, 27, 300, 2, str(tmp_path), 1), nprocs=3, join=True)
    got = [(tmp_path / f"r{r}.txt").read_text() for r in range(3)]
    assert got[1] == "own"
    assert got[0].startswith("peer:") and got[2].startswith("peer:") and "1 rank(s)" in got[0]
