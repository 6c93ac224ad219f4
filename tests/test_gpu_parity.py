"""Parity of the CUDA path (through the C ABI, via ctypes) with the oracle and the reference's
golden vectors.  Bit-exact for hashes, counts, shared-hash counts; distances / p-values within
1e-12 relative (north_star).  Needs a B200: run with `-m gpu`."""
import gzip
import json

import numpy as np
import pytest

from auriclass_b200 import _lib, synth
from auriclass_b200._lib import AC_MODE_FASTA, AC_MODE_READS
from auriclass_b200.msh import read_msh
from oracle import oracle

pytestmark = pytest.mark.gpu
D = "tests/data/"
REL = 1e-12  # relative tolerance for distance / p-value / estimates computed in double on the device


def rel_close(a, b, tol=REL):
    return a == b or abs(a - b) <= tol * max(abs(a), abs(b))


def assert_same_sketch(sk, ref):
    h, c = sk.hashes()
    assert len(h) == len(ref.hashes)
    assert np.array_equal(h, ref.hashes)
    assert np.array_equal(c, ref.counts)
    assert sk.info["n_kmers"] == ref.n_kmers
    assert sk.info["n_kept"] == ref.n_kept
    assert sk.info["total_len"] == ref.total_len
    if len(h):
        assert sk.genome_size == ref.genome_size  # same IEEE operations on both sides
        assert sk.coverage == ref.coverage


@pytest.fixture(scope="module")
def ref_msh():
    return read_msh(D + "ref_sketch.msh")


@pytest.fixture(scope="module")
def reads():
    return [oracle.read_maybe_gz(D + "NC_001416.1_1.fq.gz"), oracle.read_maybe_gz(D + "NC_001416.1_2.fq.gz")]


# ------------------------------------------------------------------ reference-pinned vectors
@pytest.mark.parametrize("fasta,idx", [("NC_001416.1.fasta.gz", 0), ("NC_001604.1.fasta.gz", 1)])
def test_fasta_sketch_equals_mash_sketch_file(engine, ref_msh, fasta, idx):
    sk = engine.sketch_fasta(D + fasta, k=27, s=50000)
    h, _ = sk.hashes()
    assert np.array_equal(h, ref_msh.references[idx].hashes)  # mash's own hash list, bit for bit
    assert sk.length == ref_msh.references[idx].length
    assert_same_sketch(sk, oracle.sketch([oracle.read_maybe_gz(D + fasta)], 27, 50000, 1))


def test_fastq_min_copies_golden(engine, ref_msh, reads):
    sk = engine.sketch_reads([D + "NC_001416.1_1.fq.gz", D + "NC_001416.1_2.fq.gz"], k=27, s=50000, m=3)
    assert_same_sketch(sk, oracle.sketch(reads, 27, 50000, 3))
    assert len(sk) == 48451 and "%g" % sk.genome_size == "48454.7"
    rows = engine.dist(engine.load_references(D + "ref_sketch.msh"), sk)
    assert [(r.common, r.denom) for r in rows] == [(48451, 48476), (0, 50000)]
    assert "%g" % rows[0].distance == "9.55405e-06" and "%g" % rows[0].pvalue == "0"
    assert rows[1].distance == 1.0 and rows[1].pvalue == 1.0


def test_workflow_methods_reproduce_reference_dicts(engine):
    # same calls and assertions as /root/reference/tests/test_correct_workflow.py:86-105 and :184-203
    from pathlib import Path

    from auriclass_b200.workflow import FastaMashWorkflow, FastqMashWorkflow
    golden = json.load(open("tests/golden/workflow_dicts.json"))

    def as_int_keys(d):
        return {k: {int(i): v for i, v in col.items()} for k, col in d.items()}

    fq = FastqMashWorkflow("test", ["tests/data/NC_001416.1_1.fq.gz", "tests/data/NC_001416.1_2.fq.gz"],
                           "tests/data/ref_sketch.msh", 27, 50_000, 3, "tests/data/clade_config.csv")
    fq.query_sketch_path = Path("tmp_data/tmpfile.msh")
    fq.sketch_fastq_query()
    fq.run_mash_dist()
    assert fq.mash_output.to_dict() == as_int_keys(golden["mash_output_to_dict_fastq"])
    assert fq.estimated_genome_size == golden["fastq_estimated_genome_size"]
    assert fq.mash_output.sort_values("Distance")["Distance"].iloc[0] == golden["fastq_minimal_distance"]

    fa = FastaMashWorkflow("test", ["tests/data/NC_001416.1.fasta.gz"], "tests/data/ref_sketch.msh", 27, 50_000, 3,
                           "tests/data/clade_config.csv")
    fa.query_sketch_path = Path("tmp_data/tmpfile.msh")
    fa.sketch_fasta_query()
    fa.run_mash_dist()
    assert fa.mash_output.to_dict() == as_int_keys(golden["mash_output_to_dict_fasta"])
    assert fa.get_error_bounds() == open("tests/golden/mash_bounds_k27_p0.99.txt").read()


def test_empty_fastq_raises_like_mash(engine):
    # /root/reference/tests/test_failing_workflow.py:51-77
    from auriclass_b200.workflow import FastqMashWorkflow
    with pytest.raises(_lib.NoRecordsError):
        engine.sketch_reads([D + "test_empty_1.fq.gz", D + "test_empty_2.fq.gz"], k=27, s=50000, m=3)
    wf = FastqMashWorkflow("t", [D + "test_empty_1.fq.gz", D + "test_empty_2.fq.gz"], D + "ref_sketch.msh", 27, 50000, 3,
                           D + "clade_config.csv")
    with pytest.raises(ValueError, match="Did not find sequence records"):
        wf.sketch_fastq_query()


# ------------------------------------------------------------------ derived vectors
@pytest.mark.parametrize("files,m,s", [((0, 1), 1, 50000), ((0, 1), 3, 1000), ((0, 1), 2, 1000), ((0, 1), 1, 1000),
                                       ((0,), 3, 50000), ((0,), 1, 50000), ((0,), 2, 1000)])
def test_min_copies_and_truncation(engine, reads, files, m, s):
    bufs = [reads[i] for i in files]
    sk = engine.sketch_bytes(bufs, k=27, s=s, m=m)
    assert_same_sketch(sk, oracle.sketch(bufs, 27, s, m))


@pytest.mark.parametrize("chunk", [1, 7, 333, 4097, 8192, 65536, 1 << 20])
def test_feed_chunking_is_invisible(engine, reads, chunk):
    if chunk < 300:
        head = reads[0][:30000]
        bufs = [head[:head.rfind(b"\n@") + 1]]  # whole records only
    else:
        bufs = reads
    sk = engine.sketch_bytes(bufs, k=27, s=1000, m=2, chunk=chunk)
    assert_same_sketch(sk, oracle.sketch(bufs, 27, 1000, 2))


@pytest.mark.parametrize("k", list(range(1, 33)))
def test_every_k(engine, k):
    fa = oracle.read_maybe_gz(D + "NC_001604.1.fasta.gz")
    sk = engine.sketch_bytes([fa], k=k, s=2000, m=1, mode=AC_MODE_FASTA)
    ref = oracle.sketch([fa], k, 2000, 1)
    assert_same_sketch(sk, ref)
    assert sk.info["hash_bits"] == (64 if k > 16 else 32)


# ------------------------------------------------------------------ edge cases of the parser
def _fasta(records, width=60, crlf=False):
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for name, seq in records:
        out.append(b">" + name + nl)
        for i in range(0, len(seq), width):
            out.append(seq[i:i + width] + nl)
    return b"".join(out)


def _rand_seq(rng, n, alphabet=b"ACGT", p=None):
    return rng.choice(list(alphabet), size=n, p=p).astype(np.uint8).tobytes()


def test_fasta_multirecord_n_lowercase_short_records(engine):
    rng = np.random.default_rng(11)
    recs = [(b"r%d desc" % i, _rand_seq(rng, int(n), b"ACGTacgtNn", [.23, .23, .23, .23, .02, .02, .01, .01, .01, .01]))
            for i, n in enumerate([5000, 26, 27, 28, 0, 1, 70000, 59, 60, 61, 12345])]
    for width in (60, 80, 1, 7, 100000):
        fa = _fasta(recs, width)
        for k, s in ((27, 500), (16, 300)):
            sk = engine.sketch_bytes([fa], k=k, s=s, m=1, mode=AC_MODE_FASTA)
            assert_same_sketch(sk, oracle.sketch([fa], k, s, 1))


def test_fasta_crlf_and_no_trailing_newline_and_blank_tail(engine):
    rng = np.random.default_rng(12)
    recs = [(b"a", _rand_seq(rng, 3000)), (b"b", _rand_seq(rng, 2000))]
    base = _fasta(recs, 70)
    for fa in (base.rstrip(b"\n"), base + b"\n\n\n", b"\n\n" + base, _fasta(recs, 70, crlf=True)):
        sk = engine.sketch_bytes([fa], k=21, s=400, m=1, mode=AC_MODE_FASTA)
        assert_same_sketch(sk, oracle.sketch([fa], 21, 400, 1))


def test_single_line_fasta_spanning_many_tiles(engine):
    rng = np.random.default_rng(13)
    fa = b">one\n" + _rand_seq(rng, 300000) + b"\n>two\n" + _rand_seq(rng, 100) + b"\n"
    sk = engine.sketch_bytes([fa], k=27, s=5000, m=1, mode=AC_MODE_FASTA, chunk=50001)
    assert_same_sketch(sk, oracle.sketch([fa], 27, 5000, 1))


def _fastq(rng, n, lens=(150,), n_rate=0.01, crlf=False, qual_at=False):
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for i in range(n):
        L = int(rng.choice(lens))
        seq = _rand_seq(rng, L, b"ACGTN", [(1 - n_rate) / 4] * 4 + [n_rate])
        qual = bytes(rng.integers(33, 74, L, dtype=np.uint8))
        if qual_at and L:
            qual = b"@" + qual[1:] if i % 3 == 0 else b"+" + qual[1:]
        out.append(b"@r%d x\n"[:-1] % i + nl + seq + nl + b"+" + nl + qual + nl)
    return b"".join(out)


def test_fastq_ragged_reads_short_reads_and_tricky_quality(engine):
    rng = np.random.default_rng(14)
    g = _rand_seq(rng, 4000)

    def reads_from_genome(n, lens):
        out = []
        for i in range(n):
            L = int(rng.choice(lens))
            st = int(rng.integers(0, len(g) - L + 1))
            q = bytes(rng.integers(33, 74, L, dtype=np.uint8))
            if L:
                q = (b"@" if i % 2 else b"+") + q[1:]  # quality lines that look like headers / separators
            out.append(b"@r%d\n" % i + g[st:st + L] + b"\n+r%d\n" % i + q + b"\n")
        return b"".join(out)

    fq = reads_from_genome(3000, (150, 151, 26, 27, 28, 1, 0, 75, 250))
    for m in (1, 2, 3):
        sk = engine.sketch_bytes([fq], k=27, s=1000, m=m)
        assert_same_sketch(sk, oracle.sketch([fq], 27, 1000, m))
    fq2 = _fastq(rng, 500, crlf=True)
    sk = engine.sketch_bytes([fq2, fq], k=27, s=1000, m=1)
    assert_same_sketch(sk, oracle.sketch([fq2, fq], 27, 1000, 1))
    # no trailing newline
    sk = engine.sketch_bytes([fq.rstrip(b"\n")], k=27, s=1000, m=2, chunk=9999)
    assert_same_sketch(sk, oracle.sketch([fq.rstrip(b"\n")], 27, 1000, 2))


def test_reads_mode_merges_fasta_and_fastq_files(engine, reads):
    fa = oracle.read_maybe_gz(D + "NC_001416.1.fasta.gz")
    bufs = [reads[0][:400000], fa, reads[1][:200000]]
    # cut at record boundaries so the FASTQ pieces stay well-formed
    bufs[0] = bufs[0][:bufs[0].rfind(b"\n@") + 1]
    bufs[2] = bufs[2][:bufs[2].rfind(b"\n@") + 1]
    sk = engine.sketch_bytes(bufs, k=27, s=3000, m=2)
    assert_same_sketch(sk, oracle.sketch(bufs, 27, 3000, 2))


def test_multiline_fastq_is_rejected_loudly(engine):
    fq = b"@r1\nACGTACGTACGTACGTACGTACGTACGTACGT\nACGTACGTACGTACGTACGTACGTACGTACGT\n+\n" + b"I" * 32 + b"\n" + b"I" * 32 + b"\n"
    with pytest.raises(_lib.AcError) as e:
        engine.sketch_bytes([fq], k=27, s=100, m=1)
    assert e.value.code == _lib.AC_E_FORMAT
    with pytest.raises(_lib.AcError):  # truncated record
        engine.sketch_bytes([b"@r1\nACGTACGTACGTACGTACGTACGTACGTACGT\n+\n"], k=27, s=100, m=1)


def test_all_records_shorter_than_k(engine):
    with pytest.raises(_lib.NoRecordsError):
        engine.sketch_bytes([b">a\nACGTACGT\n>b\nACGT\n"], k=27, s=100, m=1, mode=AC_MODE_FASTA)
    with pytest.raises(_lib.NoRecordsError):  # nothing reaches min-copies 3
        engine.sketch_bytes([b"@a\n" + b"ACGTTGCAAGGCTTAACCGGTTAGCATCGATCG" + b"\n+\n" + b"I" * 33 + b"\n"], k=27, s=100, m=3)


# ------------------------------------------------------------------ big inputs: segments, speculation, shrinking threshold
def test_multi_segment_reads_speculative_threshold(engine):
    """~110 MB of reads from a 2 Mb genome: several 32 MiB segments, tau starts at 1/16 of the hash
    space and shrinks; the oracle (all cores) must agree bit for bit."""
    contigs = synth.genome(seed=99, scale=2.0 / 12.4)
    fq = synth.fastq_reads(contigs, 340_000, seed=100)
    buf = fq.tobytes()
    assert len(buf) > 3 * (32 << 20)
    ref = oracle.sketch_parallel([buf], 27, 50000, 3)
    sk = engine.sketch_bytes([buf], k=27, s=50000, m=3, chunk=16 << 20)
    assert_same_sketch(sk, ref)
    assert sk.info["passes"] == 1 and sk.info["final_tau"] < (1 << 60)
    assert sk.info["n_candidates"] < sk.info["n_kmers"] // 8
    # device-resident feed (what bench.py's `value` times) gives the same sketch
    import torch
    t = torch.frombuffer(bytearray(buf), dtype=torch.uint8).cuda()
    import ctypes as C
    job = engine._begin(27, 50000, 42, 3, AC_MODE_READS)
    _lib.check(engine.lib.ac_sketch_feed_device(job, t.data_ptr(), t.numel(), 0, 1))
    from auriclass_b200.mash_gpu import StreamStats
    sk2 = engine._end(job, "dev", 27, 50000, StreamStats())
    assert_same_sketch(sk2, ref)


def test_many_small_segments(reads):
    """AC_SEG_MB=1: the FASTQ fixture crosses several segment boundaries (carry of the base stream,
    parser state, candidate consolidation between segments) and still matches the oracle."""
    import os
    from auriclass_b200.mash_gpu import MashGPU
    os.environ["AC_SEG_MB"] = "1"
    try:
        eng = MashGPU(0)
    finally:
        del os.environ["AC_SEG_MB"]
    try:
        contigs = synth.genome(seed=7, scale=0.01)
        fq = synth.fastq_reads(contigs, 120_000, seed=8).tobytes()  # ~38 MB: speculative, 38 segments
        sk = eng.sketch_bytes([fq], k=27, s=2000, m=3, chunk=3 << 20)
        assert_same_sketch(sk, oracle.sketch_parallel([fq], 27, 2000, 3))
        assert sk.info["passes"] == 1
        fa = synth.fasta_bytes(synth.genome(seed=9, scale=0.3))   # 3.7 MB FASTA, tau = max, 4 segments
        sk = eng.sketch_bytes([fa], k=21, s=5000, m=1, mode=AC_MODE_FASTA, chunk=999_983)
        assert_same_sketch(sk, oracle.sketch([fa], 21, 5000, 1))
        sk = eng.sketch_bytes(reads, k=27, s=50000, m=3)          # 5.2 MB of reads, tau = max
        assert_same_sketch(sk, oracle.sketch(reads, 27, 50000, 3))
    finally:
        eng.close()


def test_speculation_failure_falls_back_exactly(engine):
    """Small genome, many reads: fewer than s qualified hashes exist below the speculative
    threshold, so the pass is repeated over the retained input with tau = max."""
    rng = np.random.default_rng(5)
    contigs = [np.frombuffer(_rand_seq(rng, 30000), np.uint8)]
    fq = synth.fastq_reads(contigs, 140_000, seed=6).tobytes()  # ~45 MB > one segment
    ref = oracle.sketch_parallel([fq], 27, 50000, 3)
    sk = engine.sketch_bytes([fq], k=27, s=50000, m=3, chunk=8 << 20)
    assert_same_sketch(sk, ref)
    assert sk.info["passes"] == 2


def test_count_table_overflow_falls_back_to_sorting():
    """AC_TABLE_LOG2=10: the speculative pass's count table (1,024 slots) fills up, the pass is void and
    is repeated through the sort path with the same threshold -- same sketch, two passes."""
    import os
    from auriclass_b200.mash_gpu import MashGPU
    os.environ["AC_TABLE_LOG2"] = "10"
    try:
        eng = MashGPU(0)
    finally:
        del os.environ["AC_TABLE_LOG2"]
    try:
        contigs = synth.genome(seed=17, scale=0.02)
        fq = synth.fastq_reads(contigs, 60_000, seed=18).tobytes()   # ~19 MB: speculative
        sk = eng.sketch_bytes([fq], k=27, s=2000, m=3, chunk=5 << 20)
        assert_same_sketch(sk, oracle.sketch_parallel([fq], 27, 2000, 3))
        assert sk.info["passes"] == 2
    finally:
        eng.close()


@pytest.mark.parametrize("k,s,m", [(27, 50000, 1), (21, 10000, 1), (16, 5000, 1)])
def test_fasta_full_genome_through_count_table(engine, k, s, m):
    """C1-sized assembly (12.4 Mb, 7 contigs): large enough for the speculative threshold, so the sketch
    comes out of the count table with min_copies = 1 (k = 16: 32-bit hashes, 28-bit bucket space)."""
    fa = synth.fasta_bytes(synth.genome(seed=20240001))
    sk = engine.sketch_bytes([fa], k=k, s=s, m=m, mode=AC_MODE_FASTA)
    assert_same_sketch(sk, oracle.sketch_parallel([fa], k, s, m))
    assert sk.info["passes"] == 1 and sk.info["n_candidates"] < sk.info["n_kmers"] // 8


def test_reads_with_duplicated_low_complexity_records(engine):
    """Thousands of identical low-complexity reads: one hash occurs in every window of whole tiles, so K2's
    shared-memory staging overflows and the excess goes straight into the count table."""
    contigs = synth.genome(seed=31, scale=0.05)
    fq = synth.fastq_reads(contigs, 30_000, seed=32).tobytes()
    # a repeat unit whose k-mers include one below the speculative threshold (2^60), found with the oracle
    unit = next(u for u in (b"A", b"C", b"AC", b"AG", b"AT", b"CG", b"AAC", b"AAG", b"AAT", b"ACC", b"ACG", b"AGC", b"AACT")
                if oracle.all_hashes(b">r\n" + (u * 60)[:60] + b"\n", 27).min() < (1 << 60))
    seq = (unit * 150)[:150]
    poly = b"".join(b"@p%d\n%s\n+\n%s\n" % (i, seq, b"I" * 150) for i in range(6000))
    mono = b"".join(b"@q%d\n%s\n+\n%s\n" % (i, b"ACGT" * 37, b"I" * 148) for i in range(3000))
    cut = fq.rfind(b"\n@SYN", 0, len(fq) // 2) + 1
    buf = fq[:cut] + poly + mono + fq[cut:]
    ref = oracle.sketch_parallel([buf], 27, 3000, 3)
    sk = engine.sketch_bytes([buf], k=27, s=3000, m=3, chunk=7 << 20)
    assert_same_sketch(sk, ref)


def test_fasta_genome_scale(engine):
    contigs = synth.genome(seed=3, scale=0.25)
    fa = synth.fasta_bytes(contigs)
    sk = engine.sketch_bytes([fa], k=27, s=50000, m=1, mode=AC_MODE_FASTA)
    assert_same_sketch(sk, oracle.sketch([fa], 27, 50000, 1))


def test_sharding_invariance_property(engine, reads):
    """Order independence: any permutation / split of the read files gives the same sketch."""
    a = engine.sketch_bytes([reads[0], reads[1]], k=27, s=1000, m=3)
    b = engine.sketch_bytes([reads[1], reads[0]], k=27, s=1000, m=3)
    ha, ca = a.hashes()
    hb, cb = b.hashes()
    assert np.array_equal(ha, hb) and np.array_equal(ca, cb)


# ------------------------------------------------------------------ distance / p-value
def test_dist_matches_oracle_including_pvalues(engine):
    rng = np.random.default_rng(21)
    base = np.unique(rng.integers(0, 1 << 62, 70000, dtype=np.uint64))
    refs, lens = [], []
    for i, (n, keep) in enumerate([(50000, 1.0), (50000, 0.9), (30000, 0.5), (50000, 0.001), (10, 1.0), (50000, 0.0)]):
        mask = rng.random(base.size) < keep
        own = np.unique(rng.integers(0, 1 << 62, 70000, dtype=np.uint64))
        h = np.sort(np.concatenate([base[mask], own[: base.size - int(mask.sum())]]))[:n]
        refs.append(np.unique(h))
        lens.append(int(rng.integers(1000, 20_000_000)))
    from auriclass_b200.msh import MshReference, MshSketch
    for k in (27, 16, 21):
        m = MshSketch(kmer_size=k, sketch_size=50000, hash_bits=64 if k > 16 else 32,
                      references=[MshReference("r%d" % i, "", lens[i], refs[i] if k > 16 else np.unique(refs[i] & np.uint64(0xFFFFFFFF)))
                                  for i in range(len(refs))])
        gr = engine.load_references(m)
        for qi in (0, 1, 3, 4):
            qh = m.references[qi].hashes
            qlen = 12_400_000 if k > 16 else 48502
            q = engine.sketch_from_hashes(qh, k, 50000, qlen)
            rows = engine.dist(gr, q)
            for r, ref in zip(rows, m.references):
                d, p, c, n = oracle.dist(ref.hashes, qh, 50000, k, ref.length, qlen)
                assert (r.common, r.denom) == (c, n)
                assert rel_close(r.distance, d), (r.distance, d)
                assert rel_close(r.pvalue, p, 1e-12) or abs(r.pvalue - p) < 1e-300, (r.pvalue, p)   # north_star: 1e-12 relative
        gr.free()


def test_dist_small_s_cap(engine, ref_msh):
    lam, t7 = ref_msh.references
    from auriclass_b200.msh import MshSketch
    m = MshSketch(kmer_size=27, sketch_size=1000, references=[lam, t7])
    m.references = [type(lam)(lam.name, "", lam.length, lam.hashes[:1000]), type(t7)(t7.name, "", t7.length, t7.hashes[:1000])]
    gr = engine.load_references(m)
    q = engine.sketch_from_hashes(lam.hashes[:1000], 27, 1000, lam.length)
    rows = engine.dist(gr, q)
    for r, ref in zip(rows, m.references):
        d, p, c, n = oracle.dist(ref.hashes, lam.hashes[:1000], 1000, 27, ref.length, lam.length)
        assert (r.common, r.denom) == (c, n) and rel_close(r.distance, d)


def test_k_mismatch_is_an_error(engine, ref_msh):
    q = engine.sketch_from_hashes(ref_msh.references[0].hashes, 21, 50000, 48502)
    with pytest.raises(_lib.AcError) as e:
        engine.dist(engine.load_references(D + "ref_sketch.msh"), q)
    assert e.value.code == _lib.AC_E_PARAM_MISMATCH


def test_launch_counter_counts_our_kernels(engine, reads):
    n0 = engine.launch_count()
    engine.sketch_bytes([reads[0]], k=27, s=1000, m=1)
    assert engine.launch_count() - n0 >= 5


# ------------------------------------------------------------------ multi-GPU exchange: CUDA backend pieces
def _split_fastq(fq: bytes, parts: int):
    lines = fq.split(b"\n")[:-1]
    recs = [b"\n".join(lines[i:i + 4]) + b"\n" for i in range(0, len(lines), 4)]
    per = -(-len(recs) // parts)
    return [b"".join(recs[r * per:(r + 1) * per]) for r in range(parts)]


@pytest.mark.parametrize("P,s,m", [(2, 1000, 3), (3, 50000, 3), (4, 2000, 2)])
def test_sharded_cuda_backend_emulated_ranks(engine, reads, P, s, m):
    """P ranks emulated one after the other on one GPU: every rank's local table (ac_sketch_end_local),
    the hash-range split (ac_job_table_ranges / _export), the owner merge (ac_merge_counts) and the
    final sketch (ac_sketch_from_table) are the CUDA entry points the NCCL path uses; the exchange
    itself is done with torch indexing here.  Result must equal the oracle on all reads together."""
    import torch
    from auriclass_b200.sharded import CudaTableBackend
    be = CudaTableBackend(engine)
    fq = reads[0] + reads[1]
    shards = _split_fastq(fq, P)

    def job_for(buf):
        job = engine._begin(27, s, 42, m, AC_MODE_READS)
        a = np.frombuffer(buf, np.uint8)
        _lib.check(engine.lib.ac_sketch_feed(job, a.ctypes.data, a.size, 0, 1))
        return job

    taus = []
    for buf in shards:   # dry run: learn every rank's threshold (the all-reduce MIN of the real path)
        _, t = be.local_table(job_for(buf), lambda tau: tau, P)
        taus.append(t.tau)
    tau_cut = min(taus)
    tables = [be.local_table(job_for(buf), lambda tau: tau_cut, P) for buf in shards]
    owned_h, owned_c, keeps = [], [], []
    for r in range(P):   # all-to-all: owner r receives slice r of every rank's table
        hs, cs = [], []
        for splits, t in tables:
            o = sum(splits[:r])
            hs.append(t.hashes[o:o + splits[r]]); cs.append(t.counts[o:o + splits[r]])
        h, c = torch.cat(hs).contiguous(), torch.cat(cs).contiguous()
        n_q = be.merge_counts(h, c, m)
        owned_h.append(h[:n_q]); owned_c.append(c[:n_q]); keeps.append(n_q)
    fh, fc = torch.cat(owned_h)[:s].contiguous(), torch.cat(owned_c)[:s].contiguous()
    sk = be.build_sketch(fh, fc, 27, s, "emulated")
    ref = oracle.sketch([fq], 27, s, m)
    h, c = sk.hashes()
    assert np.array_equal(h, ref.hashes) and np.array_equal(c, ref.counts)
    assert sk.genome_size == ref.genome_size and sk.coverage == ref.coverage
    assert sum(t.counters[3] for _, t in tables) == ref.n_kmers


def test_sharded_emulated_ranks_speculative_tables(engine):
    """Shards big enough for the speculative pass: every rank's exchange table is exported from its count
    table (all keys <= tau, unqualified ones included) -- the merged sketch equals the oracle's."""
    import torch
    from auriclass_b200.sharded import CudaTableBackend
    be = CudaTableBackend(engine)
    P, s, m = 2, 2000, 3
    contigs = synth.genome(seed=41, scale=0.02)
    fq = synth.fastq_reads(contigs, 50_000, seed=42).tobytes()   # ~16 MB -> two 8 MB shards
    shards = _split_fastq(fq, P)

    def job_for(buf):
        job = engine._begin(27, s, 42, m, AC_MODE_READS)
        a = np.frombuffer(buf, np.uint8)
        _lib.check(engine.lib.ac_sketch_feed(job, a.ctypes.data, a.size, 0, 1))
        return job

    taus = [be.local_table(job_for(buf), lambda tau: tau, P)[1].tau for buf in shards]
    assert max(taus) < (1 << 60)            # speculative thresholds held
    tau_cut = min(taus)
    tables = [be.local_table(job_for(buf), lambda tau: tau_cut, P) for buf in shards]
    owned_h, owned_c = [], []
    for r in range(P):
        hs, cs = [], []
        for splits, t in tables:
            o = sum(splits[:r])
            hs.append(t.hashes[o:o + splits[r]]); cs.append(t.counts[o:o + splits[r]])
        h, c = torch.cat(hs).contiguous(), torch.cat(cs).contiguous()
        n_q = be.merge_counts(h, c, m)
        owned_h.append(h[:n_q]); owned_c.append(c[:n_q])
    fh, fc = torch.cat(owned_h)[:s].contiguous(), torch.cat(owned_c)[:s].contiguous()
    sk = be.build_sketch(fh, fc, 27, s, "emulated")
    ref = oracle.sketch_parallel([fq], 27, s, m)
    h, c = sk.hashes()
    assert np.array_equal(h, ref.hashes) and np.array_equal(c, ref.counts)
    assert sk.genome_size == ref.genome_size and sk.coverage == ref.coverage


def test_sharded_nccl_world1(engine, reads):
    """The real ShardedSketcher over an NCCL process group of one rank (all collectives run)."""
    import os
    import torch
    import torch.distributed as dist
    from auriclass_b200.sharded import ShardedSketcher
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29731")
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        sh = ShardedSketcher(engine, 0, 1)
        job = engine._begin(27, 50000, 42, 3, AC_MODE_READS)
        for i, b in enumerate(reads):
            a = np.frombuffer(b, np.uint8)
            _lib.check(engine.lib.ac_sketch_feed(job, a.ctypes.data, a.size, i, 1))
        sk = sh.finish(job, 27, 50000, 3)
        assert_same_sketch(sk, oracle.sketch(reads, 27, 50000, 3))
    finally:
        dist.destroy_process_group()


# ------------------------------------------------------------------ batch paths (BASELINE configs 4 and 5)
def _random_sketch_set(n, s, seed, bits=62):
    rng = np.random.default_rng(seed)
    base = np.unique(rng.integers(0, 1 << bits, 2 * s, dtype=np.uint64))
    out = []
    for i in range(n):
        h = base.copy()
        rep = rng.random(h.size) < rng.uniform(0, 0.3)
        h[rep] = rng.integers(0, 1 << bits, int(rep.sum()), dtype=np.uint64)
        out.append(np.unique(h)[: max(1, s - 7 * (i % 4))])        # ragged sizes
    return out


@pytest.mark.parametrize("n,s,k,bits", [(9, 1000, 27, 62), (6, 50000, 27, 64), (5, 300, 16, 32), (4, 33, 21, 62)])
def test_dist_pairs_tile_matches_oracle(engine, n, s, k, bits):
    """ac_dist_pairs (merge path, one warp per pair) against the oracle's sequential capped merge: shared-hash
    counts and denominators bit-exact, distance / p-value within 1e-12; all merge-termination branches
    (identical, disjoint, one list exhausted, cap reached) occur in the set."""
    sk = _random_sketch_set(n, s, seed=n * 100 + k, bits=min(bits, 62) if bits != 64 else 63)
    sk[1] = sk[0].copy()                                       # identical pair: d = 0, common = denom
    sk[2] = np.unique(np.random.default_rng(1).integers(1 << 62, (1 << 63) - 1, s // 2, dtype=np.uint64)) if bits > 32 \
        else np.arange(1, s // 2 + 1, dtype=np.uint64) * np.uint64(3)          # (almost) disjoint, short: exhausts early
    if bits == 32:
        sk = [np.unique(h & np.uint64(0xFFFFFFFF)) for h in sk]
    lengths = [40_000 + 1000 * i for i in range(n)]
    sset = engine.load_sketch_set(sk, lengths, k, s)
    try:
        for (i0, ni, j0, nj) in [(0, n, 0, n), (1, n - 2, 2, n - 2), (n - 1, 1, 0, 1)]:
            D, P, CM, DN = engine.dist_pairs(sset, i0, ni, sset, j0, nj)
            for i in range(ni):
                for j in range(nj):
                    d, p, cm, dn = oracle.dist(sk[i0 + i], sk[j0 + j], s, k, lengths[i0 + i], lengths[j0 + j])
                    assert (int(CM[i, j]), int(DN[i, j])) == (cm, dn), (i0 + i, j0 + j)
                    assert rel_close(D[i, j], d) and rel_close(P[i, j], p), (i0 + i, j0 + j, D[i, j], d, P[i, j], p)
        D, P, CM, DN = engine.dist_pairs(sset, 0, n, sset, 0, n)
        assert np.array_equal(CM, CM.T) and np.array_equal(DN, DN.T) and np.array_equal(D, D.T)
        assert (CM.diagonal() == DN.diagonal()).all() and (D.diagonal() == 0).all()
    finally:
        sset.free()


def _lists_from_pattern(pattern, start=1000, step=7):
    """Two ascending hash lists whose merge reads as `pattern`: 'a' / 'b' = a value only in A / only in B, 't' = a value
    in both (two merged items: the A element, its B twin right behind)."""
    a, b, v = [], [], start
    for ch in pattern:
        if ch in "at":
            a.append(v)
        if ch in "bt":
            b.append(v)
        v += step
    return np.array(a, np.uint64), np.array(b, np.uint64)


def _n_items(pattern):
    return sum(2 if ch == "t" else 1 for ch in pattern)


@pytest.mark.parametrize("block", [416, 512, 544])
def test_dist_pairs_at_block_boundaries(engine, block):
    """k5_pairs consumes the merge in blocks (416 items: 13 per lane): caps that land exactly on a block's last item, twins
    that straddle two blocks, lists that end with a block, blocks drawn from one list only -- every pattern for the
    sketch sizes S that put the cap on, one before and one behind the boundary, against the oracle's sequential merge."""
    rng = np.random.default_rng(block)
    pats = {
        "alternating": "ab" * (block + 40),
        "all_twins": "t" * (block + 40),
        "twin_straddles_block_1": "ab" * ((block - 2) // 2) + "a" + "t" + "ba" * 300,      # items block-1 / block are a twin pair
        "twin_straddles_block_2": "ab" * (block - 1) + "a" + "t" + "ab" * 200,
        "a_only_block": "a" * block + "ab" * 200,
        "b_only_block": "b" * block + "ab" * 200,
        "a_then_twins": "a" * (block - 1) + "t" * 300,
        "ends_with_block": "ab" * (block // 2),
        "ends_one_past": "ab" * (block // 2) + "a",
        "twins_end_with_block": "t" * (block // 2),
        "random_mix": "".join(rng.choice(list("abt"), size=3 * block, p=[0.35, 0.35, 0.3])),
        "random_sparse_twins": "".join(rng.choice(list("abt"), size=3 * block, p=[0.49, 0.49, 0.02])),
    }
    for name, pat in pats.items():
        assert _n_items(pat) >= block, name
    lists = {name: _lists_from_pattern(pat) for name, pat in pats.items()}
    # sketch sizes: distinct values up to the first / second block boundary of the simplest patterns, +- 1, and uncapped
    for S in sorted({block // 2 - 1, block // 2, block // 2 + 1, block - 1, block, block + 1, 2 * block - 1, 2 * block, 2 * block + 1, 10 * block}):
        sk, owner = [], []
        for name, (a, b) in lists.items():
            sk += [a, b]
            owner += [name + ":A", name + ":B"]
        lengths = [50_000 + 10 * i for i in range(len(sk))]
        sset = engine.load_sketch_set(sk, lengths, 21, S)
        try:
            n = len(sk)
            D, P, CM, DN = engine.dist_pairs(sset, 0, n, sset, 0, n)
            for i in range(n):
                for j in range(n):
                    d, p, cm, dn = oracle.dist(sk[i], sk[j], S, 21, lengths[i], lengths[j])
                    assert (int(CM[i, j]), int(DN[i, j])) == (cm, dn), (S, owner[i], owner[j], int(CM[i, j]), int(DN[i, j]), cm, dn)
                    assert rel_close(D[i, j], d) and rel_close(P[i, j], p), (S, owner[i], owner[j])
        finally:
            sset.free()


def test_dist_pairs_agrees_with_query_kernel(engine, ref_msh):
    """The tile kernel and the per-query kernel (ac_dist) give the same rows for the reference's own fixture."""
    refs = engine.load_references(D + "ref_sketch.msh")
    q = engine.sketch_from_hashes(ref_msh.references[0].hashes, 27, 50000, ref_msh.references[0].length)
    rows = engine.dist(refs, q)
    Dm, Pm, CM, DN = engine.dist_pairs(refs, 0, 2, refs, 0, 1)
    for i, r in enumerate(rows):
        assert (r.common, r.denom) == (int(CM[i, 0]), int(DN[i, 0]))
        assert rel_close(r.distance, Dm[i, 0]) and rel_close(r.pvalue, Pm[i, 0])
    q.free()


def test_classify_batch_and_all_pairs_with_the_cuda_engine(engine):
    """auriclass_b200.batch on one rank with MashGPU: rows equal the oracle's sketch + dist per query."""
    from auriclass_b200.batch import all_pairs_dense, classify_batch
    rng = np.random.default_rng(77)
    base = rng.integers(0, 4, 60_000, dtype=np.uint8)
    fastas = []
    for i in range(5):
        g = base.copy()
        mut = rng.random(g.size) < 0.003 * i
        g[mut] = (g[mut] + 1) % 4
        seq = np.frombuffer(b"ACGT", np.uint8)[g].tobytes()
        fastas.append(b">a%d\n" % i + b"\n".join(seq[o:o + 80] for o in range(0, len(seq), 80)) + b"\n")
    k, s = 27, 5000
    osk = [oracle.sketch([f], k, s, 1) for f in fastas]
    refs = engine.load_sketch_set([osk[0].hashes, osk[4].hashes], [osk[0].total_len, osk[4].total_len], k, s, names=["r0", "r4"])
    try:
        rows = classify_batch(engine, fastas, refs, k=k, s=s, dist_batch=2)
        assert len(rows) == 10
        for qi in range(5):
            for ri, rr in enumerate((0, 4)):
                d, p, cm, dn = oracle.dist(osk[rr].hashes, osk[qi].hashes, s, k, osk[rr].total_len, osk[qi].total_len)
                r = rows[2 * qi + ri]
                assert (r.reference, r.query, r.common, r.denom) == (f"r{rr}", f"query_{qi}", cm, dn)
                assert rel_close(r.distance, d) and rel_close(r.pvalue, p)
        sset = engine.load_sketch_set([o.hashes for o in osk], [o.total_len for o in osk], k, s)
        Dm, Pm, CM, DN = all_pairs_dense(engine, sset, 5, tile=2)
        for i in range(5):
            for j in range(5):
                d, p, cm, dn = oracle.dist(osk[i].hashes, osk[j].hashes, s, k, osk[i].total_len, osk[j].total_len)
                assert (int(CM[i, j]), int(DN[i, j])) == (cm, dn) and rel_close(Dm[i, j], d) and rel_close(Pm[i, j], p)
        sset.free()
    finally:
        refs.free()


# ------------------------------------------------------------------ F3 / F4: pyfastx stand-in, parallel inflate
def test_fasta_genome_size_is_a_parser_by_product(engine):
    """FastaAuriclass.parse_genome_size (classes.py:725-751) sums pyfastx `Fasta(path).size`; golden value
    48502 (/root/reference/tests/test_correct_workflow.py:197).  Here it comes from K1's base counter."""
    from pathlib import Path

    from auriclass_b200 import fastx
    from auriclass_b200.workflow import FastaMashWorkflow
    fa = FastaMashWorkflow("test", ["tests/data/NC_001416.1.fasta.gz"], "tests/data/ref_sketch.msh", 27, 50_000, 3,
                           "tests/data/clade_config.csv")
    fa.query_sketch_path = Path("tmp_data/tmpfile.msh")
    fa.sketch_fasta_query()
    fa.parse_genome_size()
    assert fa.estimated_genome_size == 48502
    fastx._sizes.clear()                                       # nobody sketched it before: Fasta.size runs the parser itself
    assert fastx.Fasta("tests/data/NC_001604.1.fasta.gz").size == 39937
    # multi-record FASTA with short records, N and lower case: size counts every sequence character
    sk = engine.sketch_bytes([b">a\nACGTNNNNacgt\nAC\n>b\nAC\n>c desc\n" + b"ACGT" * 20 + b"\n"], k=21, s=100, m=1, mode=AC_MODE_FASTA)
    assert sk.info["n_bases"] == 14 + 2 + 80 and sk.info["total_len"] == 80 and sk.info["n_records"] == 3


def test_bgzf_and_multi_member_inputs_stream_to_the_same_sketch(engine, reads, tmp_path):
    import gzip

    from auriclass_b200 import inflate
    ref = oracle.sketch(reads, 27, 50000, 3)
    p1, p2 = tmp_path / "r1.fq.gz", tmp_path / "r2.fq.gz"
    p1.write_bytes(inflate.bgzf_compress(reads[0]))                                        # parallel block inflate
    half = len(reads[1]) // 2
    p2.write_bytes(gzip.compress(reads[1][:half], 1) + gzip.compress(reads[1][half:], 1))    # members cut mid-record
    sk = engine.sketch_reads([p1, p2], k=27, s=50000, m=3)
    assert_same_sketch(sk, ref)
    assert sk.stats.inflate_s > 0 and sk.stats.bytes_in == len(reads[0]) + len(reads[1])


def test_merge_counts_exact_for_clustered_keys_and_both_code_paths(engine):
    """ac_merge_counts (owner merge of the multi-GPU exchange): sums of equal hashes, >= m kept, ascending.
    Keys crowded into a few sort buckets (shared top bits) exercise the bucket sort's per-bucket heap sort; the
    all-ones key cannot live in the count table and forces the radix-sort fallback.  Both must be exact."""
    import torch
    from auriclass_b200.sharded import CudaTableBackend
    be = CudaTableBackend(engine)
    rng = np.random.default_rng(123)
    for case in ("uniform", "crowded", "all_ones"):
        if case == "uniform":
            base = rng.integers(0, 1 << 57, 20_000, dtype=np.uint64)
        elif case == "crowded":   # 3 clusters of nearby values: thousands of keys per bucket
            base = np.concatenate([(np.uint64(c) << np.uint64(50)) + rng.integers(0, 1 << 20, 3000, dtype=np.uint64) for c in (1, 7, 40)])
        else:
            base = np.concatenate([rng.integers(0, 1 << 63, 5000, dtype=np.uint64), np.array([2**64 - 1] * 3, np.uint64)])
        keys = np.concatenate([base, base[::2], base[::3], rng.integers(0, 1 << 57, 1000, dtype=np.uint64)])
        cnts = rng.integers(1, 3, keys.size).astype(np.uint32)
        perm = rng.permutation(keys.size)
        keys, cnts = keys[perm], cnts[perm]
        for m in (1, 3):
            h = torch.from_numpy(keys.view(np.int64).copy()).cuda()
            c = torch.from_numpy(cnts.view(np.int32).copy()).cuda()
            n_q = be.merge_counts(h, c, m)
            torch.cuda.synchronize()
            u, inv = np.unique(keys, return_inverse=True)
            tot = np.bincount(inv, weights=cnts).astype(np.uint64)
            want_h, want_c = u[tot >= m], tot[tot >= m]
            assert n_q == want_h.size, (case, m)
            got_h = h[:n_q].cpu().numpy().view(np.uint64)
            got_c = c[:n_q].cpu().numpy().view(np.uint32)
            assert np.array_equal(got_h, want_h) and np.array_equal(got_c.astype(np.uint64), want_c), (case, m)


@pytest.mark.parametrize("k,m", [(12, 2), (16, 3), (17, 1), (24, 2), (32, 3)])
def test_speculative_count_table_for_other_k(engine, k, m):
    """The count-table pass with 32-bit hashes (k <= 16: 27-bit bucket space), the first 64-bit k, and the widest k:
    ~10 MB of reads is enough for the speculative threshold at s = 1000."""
    contigs = synth.genome(seed=50 + k, scale=0.02)
    fq = synth.fastq_reads(contigs, 32_000, seed=60 + k).tobytes()
    assert len(fq) > (8 << 20)
    ref = oracle.sketch_parallel([fq], k, 1000, m)
    sk = engine.sketch_bytes([fq], k=k, s=1000, m=m, chunk=3 << 20)
    assert_same_sketch(sk, ref)
    assert sk.info["passes"] == 1 and sk.info["n_candidates"] < sk.info["n_kmers"] // 8
    assert sk.info["hash_bits"] == (64 if k > 16 else 32)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [28, 27, 16])
def test_flat_loop_redo_list_overflow(engine, k):
    """K2's flat loop parks windows whose strands tie on the first four bases in a 128-entry list per 2,048 windows.
    (AT)n reads tie on EVERY window for even k ((AATT)n on every other one), so whole sub-tiles overflow the list
    and the excess is re-hashed on the spot; the sketch must not notice.  Count-table pass, ~10 MB of reads."""
    contigs = synth.genome(seed=70 + k, scale=0.02)
    fq = synth.fastq_reads(contigs, 32_000, seed=80 + k).tobytes()
    ties = b"".join(b"@t%d\n%s\n+\n%s\n" % (i, (u * 80)[:150], b"I" * 150)
                    for i in range(4000) for u in (b"AT", b"AATT", b"ACGT"))
    cut = fq.rfind(b"\n@SYN", 0, len(fq) // 2) + 1
    buf = fq[:cut] + ties + fq[cut:]
    assert len(buf) > (8 << 20)
    ref = oracle.sketch_parallel([buf], k, 1000, 2)
    sk = engine.sketch_bytes([buf], k=k, s=1000, m=2, chunk=3 << 20)
    assert_same_sketch(sk, ref)
    assert sk.info["passes"] == 1


def test_fastq_many_small_segments_device_and_host_fed(reads):
    """Many small segments (AC_SEG_MB=1: the parser state, the stream carry and the look-back status words are handed from
    segment to segment 38 times), device-resident and host-fed, two files merged, a pass that is repeated after a failed
    speculation, and the tau = max path: every sketch must equal the oracle's."""
    import os
    import torch
    from auriclass_b200.mash_gpu import MashGPU, StreamStats
    os.environ["AC_SEG_MB"] = "1"
    try:
        eng = MashGPU(0)
    finally:
        del os.environ["AC_SEG_MB"]
    try:
        contigs = synth.genome(seed=17, scale=0.01)
        fq = synth.fastq_reads(contigs, 120_000, seed=18).tobytes()   # ~38 MB, 38 segments
        ref = oracle.sketch_parallel([fq], 27, 2000, 3)
        t = torch.frombuffer(bytearray(fq), dtype=torch.uint8).cuda()
        for _ in range(2):   # device-resident feed
            job = eng._begin(27, 2000, 42, 3, AC_MODE_READS)
            _lib.check(eng.lib.ac_sketch_feed_device(job, t.data_ptr(), t.numel(), 0, 1))
            assert_same_sketch(eng._end(job, "dev", 27, 2000, StreamStats()), ref)
        assert_same_sketch(eng.sketch_bytes([fq], k=27, s=2000, m=3, chunk=3 << 20), ref)       # host-fed in 3 MB chunks
        half = fq.rfind(b"\n@SYN", 0, len(fq) // 2) + 1
        two = [fq[:half], fq[half:]]                                                           # two files, one sketch
        assert_same_sketch(eng.sketch_bytes(two, k=27, s=2000, m=3, chunk=5 << 20), oracle.sketch_parallel(two, 27, 2000, 3))
        # too few qualified hashes below 1/32 of the space: the pass is repeated with tau = max (look-ahead restarts)
        low = synth.fastq_reads(contigs, 30_000, seed=19).tobytes()
        sk = eng.sketch_bytes([low], k=27, s=50000, m=3, chunk=2 << 20)
        assert_same_sketch(sk, oracle.sketch_parallel([low], 27, 50000, 3))
        sk = eng.sketch_bytes(reads, k=27, s=50000, m=3)                                        # 5.2 MB, tau = max
        assert_same_sketch(sk, oracle.sketch(reads, 27, 50000, 3))
    finally:
        eng.close()


def test_fastq_long_reads_spanning_tiles_and_subtiles(engine):
    """Long-read FASTQ: sequence and quality lines of 3 kb .. 120 kb run across many 8 KiB parser tiles and
    2,048-window hash sub-tiles (a header may sit anywhere inside a tile, a whole tile may be one line class).
    Small input (tau = max, branchy loop) and a 12 MB one (count table, flat loop), both against the oracle."""
    rng = np.random.default_rng(21)
    g = _rand_seq(rng, 300_000)

    def long_reads(n, lo, hi, n_rate=0.001):
        out = []
        for i in range(n):
            L = int(rng.integers(lo, hi))
            st = int(rng.integers(0, len(g) - L + 1))
            seq = bytearray(g[st:st + L])
            for p in rng.integers(0, L, int(L * n_rate)):
                seq[int(p)] = ord("N")
            q = bytes(rng.integers(33, 74, L, dtype=np.uint8))
            q = (b"@" if i % 2 else b"+") + q[1:]
            out.append(b"@lr%d runid=x\n" % i + bytes(seq) + b"\n+\n" + q + b"\n")
        return b"".join(out)

    small = long_reads(12, 3_000, 120_000)
    assert len(small) < (4 << 20)
    sk = engine.sketch_bytes([small], k=27, s=1000, m=1, chunk=100_003)
    assert_same_sketch(sk, oracle.sketch([small], 27, 1000, 1))
    big = long_reads(110, 3_000, 120_000)
    assert len(big) > (8 << 20)
    for k, m in ((27, 2), (16, 1)):
        sk = engine.sketch_bytes([big], k=k, s=1000, m=m, chunk=3 << 20)
        assert_same_sketch(sk, oracle.sketch_parallel([big], k, 1000, m))
        assert sk.info["passes"] == 1


# ------------------------------------------------------------------ round-2 additions
def test_device_fed_default_size_segments_match_oracle(engine):
    """The benchmarked configuration's shape: a device-resident feed that spans three DEFAULT-size (256 MiB) segments
    (two full ones and a tail), count-table pass with tau updates between segments, against the oracle."""
    import torch
    n_reads = 1_750_000                                   # 563 MB of 4-line FASTQ
    contigs = synth.genome(seed=20240001, scale=0.25)    # 3.1 Mb -> ~85x
    text = synth.fastq_reads(contigs, n_reads, seed=77)
    assert text.size > 2 * (256 << 20)
    dev = torch.from_numpy(text).cuda()
    job = engine._begin(27, 50000, 42, 3, AC_MODE_READS)
    _lib.check(engine.lib.ac_sketch_feed_device(job, dev.data_ptr(), text.size, 0, 1))
    from auriclass_b200.mash_gpu import StreamStats
    sk = engine._end(job, "dev", 27, 50000, StreamStats())
    ref = oracle.sketch_parallel([text], 27, 50000, 3)
    assert sk.info["passes"] == 1
    assert_same_sketch(sk, ref)
    assert len(sk) == 50000
    sk.free()
    # the same bytes from host memory in 32 MiB feeds (the e2e arm of bench.py) give the same sketch
    job = engine._begin(27, 50000, 42, 3, AC_MODE_READS)
    step = 32 << 20
    for off in range(0, text.size, step):
        n = min(step, text.size - off)
        _lib.check(engine.lib.ac_sketch_feed(job, text[off:].ctypes.data, n, 0, 1 if off + n >= text.size else 0))
    sk = engine._end(job, "host", 27, 50000, StreamStats())
    assert_same_sketch(sk, ref)
    sk.free()


def test_check_dependencies_passes_with_a_device_present():
    """F2: the replacement of the `mash -h` probe (general.py:181-205) on a box that has the library and a B200."""
    from auriclass_b200.workflow import check_dependencies
    assert check_dependencies() is None
    lib = _lib.load()
    assert lib.ac_probe() == 0 and lib.ac_device_count() >= 1


def test_sharded_public_entry_world1_and_workflow(engine, reads, tmp_path):
    """MashGPU.sketch_reads_sharded (read-block cutter -> feeds -> exchange) over an NCCL group of one rank, with a
    block size small enough that the two fixture files fall into ~40 blocks; then the same through the reference-shaped
    workflow method, which switches to the sharded entry only when the group has more than one rank."""
    import os
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29733")
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        paths = [D + "NC_001416.1_1.fq.gz", D + "NC_001416.1_2.fq.gz"]
        sk = engine.sketch_reads_sharded(paths, k=27, s=50000, m=3, block_bytes=64 << 10)
        ref = oracle.sketch(reads, 27, 50000, 3)
        assert_same_sketch(sk, ref)
        assert sk.name == paths[0] and sk.stats.bytes_in == sum(len(r) for r in reads)
        from auriclass_b200 import workflow
        assert workflow.process_group_world() == 1
        with pytest.raises(_lib.AcError):     # a rank whose blocks are not FASTQ: the error surfaces, nothing hangs
            p = tmp_path / "bad.fq"
            p.write_bytes(b"@r1\nACGT\nIIII\n" * 10)
            engine.sketch_reads_sharded([str(p)], k=27, s=100, m=1)
    finally:
        for sh in engine._sharders.values():
            sh.close()
        engine._sharders.clear()
        dist.destroy_process_group()


from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

_base = st.sampled_from("ACGTACGTACGTACGTNnacgtRY-")


@st.composite
def _ragged_fastq(draw):
    """Random reads (SURVEY.md section 4 'the new build must add'): N, lower case, IUPAC codes, records shorter than k,
    CRLF or LF line ends, an optional missing final newline; a few reads are repeated so that m > 1 keeps something."""
    n = draw(st.integers(1, 40))
    eol = draw(st.sampled_from(["\n", "\r\n"]))
    recs = []
    for i in range(n):
        L = draw(st.integers(1, 90))
        seq = "".join(draw(st.lists(_base, min_size=L, max_size=L)))
        rec = f"@r{i} x{eol}{seq}{eol}+{eol}{'I' * L}{eol}"
        recs.extend([rec] * draw(st.sampled_from([1, 1, 2, 3])))
    txt = "".join(recs)
    if draw(st.booleans()):
        txt = txt[: -len(eol)]
    return txt.encode()


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(fq=_ragged_fastq(), k=st.sampled_from([5, 16, 21, 27, 32]), s=st.sampled_from([3, 50, 10000]), m=st.sampled_from([1, 2, 3]),
       chunk=st.sampled_from([None, 7, 64, 1000]))
def test_hypothesis_random_reads_match_oracle(engine, fq, k, s, m, chunk):
    try:
        ref = oracle.sketch([fq], k, s, m)
    except ValueError:
        ref = None
    if ref is None or len(ref.hashes) == 0:
        with pytest.raises(_lib.AcError):
            engine.sketch_bytes([fq], k=k, s=s, m=m, chunk=chunk)
        return
    sk = engine.sketch_bytes([fq], k=k, s=s, m=m, chunk=chunk)
    assert_same_sketch(sk, ref)
    sk.free()


# ------------------------------------------------------------------ peer-memory exchange (csrc/ac_xchg.cu)
def _peer_group(P, s_cap=65536, inbox=0):
    """P ranks emulated in ONE process on one GPU: a context + window per rank, wired together directly
    (ac_xchg_set_peer_local) -- the kernels, flags and epochs are those of the multi-process path, only the pointers
    do not come from CUDA IPC."""
    from auriclass_b200.mash_gpu import MashGPU
    from auriclass_b200.sharded import PeerWindow
    engines = [MashGPU(0) for _ in range(P)]
    wins = [PeerWindow(e, r, P, s_cap, inbox) for r, e in enumerate(engines)]
    for a in wins:
        for r, b in enumerate(wins):
            if a is not b:
                a.set_peer_local(r, b)
    return engines, wins


def _close_group(engines, wins):
    for w in wins:
        w.close()
    for e in engines:
        e.close()


def _feed_job(eng, bufs, k, s, m):
    job = eng._begin(k, s, 42, m, AC_MODE_READS)
    for i, b in enumerate(bufs):
        a = np.frombuffer(b, np.uint8)
        _lib.check(eng.lib.ac_sketch_feed(job, a.ctypes.data if a.size else None, a.size, i, 1))
    return job


def _exchange(engines, wins, shards, k, s, m):
    """launch every rank's phase before any rank's next phase is waited for (one host thread plays all ranks)"""
    jobs = [_feed_job(e, sh, k, s, m) for e, sh in zip(engines, shards)]
    for w, j in zip(wins, jobs):
        w.begin(j, k, s, m)
    for w in wins:
        w.merge()
    return [w.end("reads", k, s) for w in wins]


@pytest.mark.parametrize("P,k,s,m", [(2, 27, 1000, 3), (3, 27, 50000, 3), (4, 21, 2000, 2), (2, 16, 700, 1), (1, 27, 50000, 3)])
def test_peer_exchange_emulated_ranks_small_tables(reads, P, k, s, m):
    """sort-mode local tables (inputs below the speculation size); the same windows carry three exchanges in a row
    (epochs), the second with an empty rank"""
    engines, wins = _peer_group(P)
    try:
        fq = reads[0] + reads[1]
        ref = oracle.sketch([fq], k, s, m)
        for trial in range(3):
            shards = [[x] for x in _split_fastq(fq, P)]
            if trial == 1 and P > 1:
                shards = [[b"".join(x[0] for x in shards)]] + [[b""] for _ in range(P - 1)]
            sks = _exchange(engines, wins, shards, k, s, m)
            for sk in sks:
                h, c = sk.hashes()
                assert np.array_equal(h, ref.hashes) and np.array_equal(c, ref.counts)
                assert sk.genome_size == ref.genome_size and sk.coverage == ref.coverage
                assert sk.info["n_kmers"] == ref.n_kmers and sk.info["n_kept"] == ref.n_kept
                sk.free()
            assert all(w.last_stats["tau_cut"] == wins[0].last_stats["tau_cut"] for w in wins)
    finally:
        _close_group(engines, wins)


def test_peer_exchange_emulated_ranks_count_tables_and_dist():
    """count-table local passes (speculative thresholds): 3 ranks x 8 MB shards; the merged sketch feeds dist"""
    P, s, m = 3, 3000, 3
    contigs = synth.genome(seed=43, scale=0.02)
    fq = synth.fastq_reads(contigs, 75_000, seed=44).tobytes()
    engines, wins = _peer_group(P)
    try:
        sks = _exchange(engines, wins, [[x] for x in _split_fastq(fq, P)], 27, s, m)
        ref = oracle.sketch_parallel([fq], 27, s, m)
        assert wins[0].last_stats["tau_cut"] < (1 << 60)
        for sk in sks:
            h, c = sk.hashes()
            assert np.array_equal(h, ref.hashes) and np.array_equal(c, ref.counts)
            assert sk.genome_size == ref.genome_size and sk.coverage == ref.coverage
        fa = synth.fasta_bytes(contigs)
        rsk = engines[0].sketch_bytes([fa], k=27, s=s, m=1, mode=AC_MODE_FASTA)
        gr = engines[0].load_sketch_set([rsk.hashes()[0]], [rsk.length], 27, s, names=["g"])
        row = engines[0].dist(gr, sks[0])[0]
        d, p, cm, dn = oracle.dist(rsk.hashes()[0], ref.hashes, s, 27, rsk.length, int(ref.genome_size))
        assert (row.common, row.denom) == (cm, dn) and rel_close(row.distance, d)
        gr.free()
    finally:
        _close_group(engines, wins)


def test_peer_exchange_failures_reach_every_rank():
    """a rank with malformed reads, and a rank that never fed anything (failed = 1): every rank gets an error, the
    failing rank its own; then the same windows complete a good exchange (nothing is left half-posted).  A share
    beyond the inbox capacity is reported, not truncated."""
    P, k, s, m = 3, 27, 500, 1
    engines, wins = _peer_group(P, inbox=1024)
    try:
        contigs = synth.genome(seed=9, scale=0.0005)
        fq = synth.fastq_reads(contigs, 300, seed=10).tobytes()
        shards = _split_fastq(fq, P)
        bad = b"@r1\nACGTACGTACGTACGTACGTACGTACGTACGT\nIIII\n" * 5       # no '+' line
        jobs = [_feed_job(e, [sh], k, s, m) for e, sh in zip(engines, [shards[0], bad, shards[2]])]
        for w, j in zip(wins, jobs):
            w.begin(j, k, s, m)
        for w in wins:
            w.merge()
        errs = []
        for w in wins:
            with pytest.raises(_lib.AcError) as e:
                w.end("x", k, s)
            errs.append(e.value.code)
        assert errs[1] == _lib.AC_E_FORMAT and errs[0] == errs[2] == _lib.AC_E_PEER
        jobs = [_feed_job(e, [sh], k, s, m) for e, sh in zip(engines, shards)]
        for r, (w, j) in enumerate(zip(wins, jobs)):
            if r == 2:
                engines[2].lib.ac_sketch_abort(j)
                w.begin(None, k, s, m, failed=True)
            else:
                w.begin(j, k, s, m)
        for w in wins:
            w.merge()
        for w in wins:
            with pytest.raises(_lib.AcError):
                w.end("x", k, s)
        sks = _exchange(engines, wins, [[x] for x in shards], k, s, m)
        ref = oracle.sketch([fq], k, s, m)
        for sk in sks:
            assert np.array_equal(sk.hashes()[0], ref.hashes)
        # overflow: ~45,000 distinct k-mers per rank against 1,024-entry inboxes
        big = synth.fastq_reads(synth.genome(seed=11, scale=0.004), 1000, seed=12).tobytes()
        jobs = [_feed_job(e, [sh], k, 60000, m) for e, sh in zip(engines, _split_fastq(big, P))]
        for w, j in zip(wins, jobs):
            w.begin(j, k, 60000, m)
        for w in wins:
            w.merge()
        for w in wins:
            with pytest.raises(_lib.AcError) as e:
                w.end("x", k, 60000)
            assert e.value.code in (_lib.AC_E_NOMEM, _lib.AC_E_PEER)
    finally:
        _close_group(engines, wins)


def test_peer_exchange_through_the_sharder_world1(engine, reads):
    """ShardedSketcher picks the peer transport on a CUDA backend: NCCL group of one rank, IPC handle exchange included"""
    import os
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29735")
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        sh = engine.sharder(None)
        assert "peer memory" in sh.transport
        sk = sh.finish(_feed_job(engine, reads, 27, 50000, 3), 27, 50000, 3, name="reads")
        assert sh.peer is not None and len(sh.peer.handle_bytes()) == 64
        assert_same_sketch(sk, oracle.sketch(reads, 27, 50000, 3))
        sk2 = sh.finish(_feed_job(engine, reads[:1], 27, 100000, 2), 27, 100000, 2)     # outgrows the window: rebuilt
        assert sh.peer.sketch_cap >= 100000
        assert_same_sketch(sk2, oracle.sketch(reads[:1], 27, 100000, 2))
    finally:
        for s_ in engine._sharders.values():
            s_.close()
        engine._sharders.clear()
        dist.destroy_process_group()


def test_fastq_phase_guess_from_the_text_is_verified(engine):
    """k1q_fused reads the line phase of a tile off the text (a line that starts with '@' whose next-but-one line starts with
    '+' is taken for a header) and checks it against the chained value after the emit side.  Records whose SEQUENCE line
    starts with '@' and whose quality line starts with '+' make that rule misfire on half of the tiles; the stream must not
    depend on it: same sketch for every shift of the tile grid (a prefix of k-mer-free records of varying size), and equal
    to the oracle's on the same reads with the stray '@' spelled 'N' (both are just invalid bases to a strict 4-line parser)."""
    rng = np.random.default_rng(99)
    g = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 30_000)].tobytes()
    recs_gpu, recs_ora = [], []
    for i in range(6000):
        L = int(rng.integers(40, 120))
        o = int(rng.integers(0, len(g) - L))
        seq = g[o:o + L]
        hdr = b"@r%d" % i
        recs_gpu.append(hdr + b"\n@" + seq + b"\n+\n+" + b"I" * L + b"\n")
        recs_ora.append(hdr + b"\nN" + seq + b"\n+\n+" + b"I" * L + b"\n")
    body_gpu, body_ora = b"".join(recs_gpu), b"".join(recs_ora)
    assert len(body_gpu) > 40 * 16384
    ref = oracle.sketch([body_ora], 21, 2000, 2)
    assert len(ref.hashes) == 2000
    for shift in (0, 1, 2, 3, 5, 64, 777, 16383):
        prefix = b"@p\n" + b"N" * shift + b"\n+\n" + b"I" * shift + b"\n" if shift else b""
        sk = engine.sketch_bytes([prefix + body_gpu], k=21, s=2000, m=2)
        h, c = sk.hashes()
        assert np.array_equal(h, ref.hashes) and np.array_equal(c, ref.counts), shift
        assert sk.info["n_kmers"] == ref.n_kmers
        sk.free()


def test_peer_exchange_owner_merge_falls_back_to_sorting(reads):
    """the owner merge runs in a window of the count table; inboxes that do not fit it (here: a 1,024-slot window against
    ~48,000 entries per owner) raise the overflow flag and the merge is redone by sorting -- same sketch"""
    import os
    os.environ["AC_XCHG_MERGE_LOG2"] = "10"
    engines, wins = _peer_group(2)
    try:
        fq = reads[0] + reads[1]
        sks = _exchange(engines, wins, [[x] for x in _split_fastq(fq, 2)], 27, 50000, 2)
        ref = oracle.sketch([fq], 27, 50000, 2)
        for sk in sks:
            h, c = sk.hashes()
            assert np.array_equal(h, ref.hashes) and np.array_equal(c, ref.counts)
            assert sk.genome_size == ref.genome_size and sk.coverage == ref.coverage
    finally:
        del os.environ["AC_XCHG_MERGE_LOG2"]
        _close_group(engines, wins)
