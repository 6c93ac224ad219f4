"""The oracle (oracle/mash_oracle.c) against every golden vector the reference's own tests
hold for this path plus the derived vectors of App. C.2/C.4.
Runs on CPU."""
import numpy as np
import pytest

from auriclass_b200.msh import read_msh
from oracle import oracle

D = "tests/data/"
ERROR_BOUNDS_TEXT_REF = open("tests/golden/mash_bounds_k27_p0.99.txt").read()


@pytest.fixture(scope="module")
def ref_msh():
    return read_msh(D + "ref_sketch.msh")


@pytest.fixture(scope="module")
def reads():
    return [oracle.read_maybe_gz(D + "NC_001416.1_1.fq.gz"), oracle.read_maybe_gz(D + "NC_001416.1_2.fq.gz")]


def test_murmur_known_answer():
    # SURVEY.md section 8c check value
    assert oracle.murmur3_x64_128(b"GGGCGGCGACCTCGCGGGTTTTCGCTA") == (0xA0F4A64133977590, 0x8902DD01600CB84C)


@pytest.mark.parametrize("fasta,idx,n", [("NC_001416.1.fasta.gz", 0, 48476), ("NC_001604.1.fasta.gz", 1, 39770)])
def test_fasta_sketch_equals_reference_msh(ref_msh, fasta, idx, n):
    # tests/data/ref_sketch.msh holds mash's own sorted hash lists of the two phage genomes
    sk = oracle.sketch([oracle.read_maybe_gz(D + fasta)], 27, 50000, 1)
    assert len(sk.hashes) == n
    assert np.array_equal(sk.hashes, ref_msh.references[idx].hashes)
    assert sk.total_len == ref_msh.references[idx].length


def test_fastq_min_copies_golden(ref_msh, reads):
    # /root/reference/tests/test_correct_workflow.py:18-25, :99, :105
    sk = oracle.sketch(reads, 27, 50000, 3)
    assert len(sk.hashes) == 48451
    assert "%g" % sk.genome_size == "48454.7"
    lam, t7 = ref_msh.references
    d, p, c, n = oracle.dist(lam.hashes, sk.hashes, 50000, 27, lam.length, int(sk.genome_size))
    assert (c, n) == (48451, 48476) and "%g" % d == "9.55405e-06" and "%g" % p == "0"
    d, p, c, n = oracle.dist(t7.hashes, sk.hashes, 50000, 27, t7.length, int(sk.genome_size))
    assert (c, n) == (0, 50000) and d == 1.0 and p == 1.0


def test_fasta_query_golden(ref_msh):
    # test_correct_workflow.py:28-38: 48476/48476 distance 0 p 0; 0/50000 distance 1 p 1
    lam, t7 = ref_msh.references
    d, p, c, n = oracle.dist(lam.hashes, lam.hashes, 50000, 27, lam.length, lam.length)
    assert (c, n, d) == (48476, 48476, 0.0) and "%g" % p == "0"
    d, p, c, n = oracle.dist(t7.hashes, lam.hashes, 50000, 27, t7.length, lam.length)
    assert (c, n, d, p) == (0, 50000, 1.0, 1.0)


def test_bounds_text_golden():
    # test_correct_workflow.py:26
    assert oracle.bounds_text(27, 0.99) == ERROR_BOUNDS_TEXT_REF


# SURVEY.md App. C.2 (derived by the validated restatement)
C2 = [
    ((0, 1), 3, 50000, 48451, 0xFFFB03E04778B43C, "48454.7", "48451/48476", "9.55405e-06"),
    ((0, 1), 1, 50000, 50000, 0x5308C256352F4743, "154153", "15869/50000", "0.0270427"),
    ((0, 1), 3, 1000, 1000, 0x053716FE1D3B99F0, "49087.3", "998/1000", "3.70927e-05"),
    ((0, 1), 2, 1000, 1000, 0x051DBBBF74569EF5, "50037.7", "976/1000", "0.000452597"),
    ((0, 1), 1, 1000, 1000, 0x01B008775512BCBD, "151692", "305/1000", "0.0281666"),
    ((0,), 3, 50000, 48425, 0xFFFB03E04778B43C, "48428.7", "48425/48476", "1.94981e-05"),
    ((0,), 1, 50000, 50000, 0x89AA368FBDAEB15F, "92979.4", "26098/50000", "0.0139636"),
    ((0,), 2, 1000, 1000, 0x0536EC12F3B33730, "49093.5", "997/1000", "5.56808e-05"),
]


@pytest.mark.parametrize("files,m,s,n,mx,gs,match,dist", C2)
def test_derived_vectors(ref_msh, reads, files, m, s, n, mx, gs, match, dist):
    sk = oracle.sketch([reads[i] for i in files], 27, s, m)
    assert len(sk.hashes) == n and int(sk.hashes[-1]) == mx and "%g" % sk.genome_size == gs
    lam = ref_msh.references[0]
    d, p, c, dn = oracle.dist(lam.hashes, sk.hashes, min(s, 50000), 27, lam.length, int(sk.genome_size))
    assert f"{c}/{dn}" == match and "%g" % d == dist


def test_streaming_heap_equals_exact_counting(reads):
    # App. C.3: Mash's streaming MinHashHeap yields the same hash SET as exact counting
    for m, s in [(3, 50000), (3, 1000), (2, 1000), (1, 1000)]:
        a = oracle.sketch(reads, 27, s, m)
        b = oracle.sketch_stream(reads, 27, s, m)
        assert np.array_equal(a.hashes, b.hashes)
    c = oracle.sketch_parallel(reads, 27, 1000, 3)
    assert np.array_equal(c.hashes, oracle.sketch(reads, 27, 1000, 3).hashes)


def test_small_vectors_n_and_lowercase():
    # App. C.4
    lam60 = b"GGGCGGCGACCTCGCGGGTTTTCGCTATTTATGAAAATTTTCCGGTTTAAGGCGTTTCCG"
    h = oracle.all_hashes(b">x\n" + lam60 + b"\n", 27)
    assert [int(x) for x in h[:3]] == [0xA0F4A64133977590, 0x0219E621B7975B36, 0x372CB1C9425EE8AA]
    withn = bytearray(lam60)
    withn[30] = ord("N")
    want = [0xA0F4A64133977590, 0x0219E621B7975B36, 0x372CB1C9425EE8AA, 0x2C7620841CA11562,
            0x23E32A90685FD4F4, 0xC27B5D183A4B18C1, 0xC1ABF5D844A975D8]
    assert [int(x) for x in oracle.all_hashes(b">x\n" + bytes(withn) + b"\n", 27)] == want
    assert [int(x) for x in oracle.all_hashes(b">x\n" + bytes(withn).lower() + b"\n", 27)] == want


def test_k21_and_32bit_mode():
    fa = oracle.read_maybe_gz(D + "NC_001416.1.fasta.gz")
    h21 = oracle.all_hashes(fa, 21)
    assert len(h21) == 48482 and int(h21.min()) == 0x0000D54406B440FF and int(h21.max()) == 0xFFFD816AA0824146
    h16 = oracle.all_hashes(fa, 16)
    assert [int(x) for x in h16[:3]] == [0x5202F588, 0xA3ECF14F, 0x816D7457]
    assert len(h16) == 48487 and len(np.unique(h16)) == 48485


def test_empty_fastq_has_no_records():
    # /root/reference/tests/test_failing_workflow.py:51-77
    sk = oracle.sketch([oracle.read_maybe_gz(D + "test_empty_1.fq.gz"), oracle.read_maybe_gz(D + "test_empty_2.fq.gz")], 27, 50000, 3)
    assert len(sk.hashes) == 0 and sk.n_records == 0
