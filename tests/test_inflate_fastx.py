"""Host-side duties around the hot path (no GPU): parallel / pipelined gzip inflate
and the pyfastx stand-in's format sniffing (F3; `Fasta.size` needs the GPU parser and is covered in
tests/test_gpu_parity.py)."""
import gzip
import sys
import zlib
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from auriclass_b200 import fastx, inflate  # noqa: E402
from auriclass_b200.mash_gpu import StreamStats  # noqa: E402


def _text(n, seed=0):
    rng = np.random.default_rng(seed)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].tobytes()
    return b"".join(b"@r%d\n%s\n+\n%s\n" % (i, seq[o:o + 150], b"I" * len(seq[o:o + 150])) for i, o in enumerate(range(0, n, 150)))


@pytest.mark.parametrize("kind", ["plain", "gzip", "multi", "bgzf", "bgzf_empty", "empty_plain"])
@pytest.mark.parametrize("threads", [1, 4])
def test_iter_inflated_returns_the_exact_bytes(tmp_path, kind, threads):
    data = _text(3_000_000, seed=3)
    p = tmp_path / "x"
    if kind == "plain":
        p.write_bytes(data)
    elif kind == "gzip":
        p.write_bytes(gzip.compress(data, 4))
    elif kind == "multi":                    # concatenated members, cut at arbitrary bytes (legal gzip)
        cuts = [0, 1, 700_001, 700_002, 2_000_000, len(data)]
        p.write_bytes(b"".join(gzip.compress(data[a:b], 1) for a, b in zip(cuts, cuts[1:])))
    elif kind == "bgzf":
        p.write_bytes(inflate.bgzf_compress(data))
    elif kind == "bgzf_empty":
        data = b""
        p.write_bytes(inflate.bgzf_compress(data))
    else:
        data = b""
        p.write_bytes(b"")
    st = StreamStats()
    got = b"".join(inflate.iter_inflated(p, st, threads))
    assert got == data
    if kind in ("gzip", "multi", "bgzf"):
        assert st.inflate_s > 0


def test_bgzf_is_what_gzip_readers_accept_and_block_table_is_exact(tmp_path):
    data = _text(1_000_000, seed=5)
    bz = inflate.bgzf_compress(data)
    assert gzip.decompress(bz) == data                       # BGZF is valid multi-member gzip
    table = inflate.bgzf_block_table(bz)
    assert table is not None and sum(size for _, size, _ in table) == len(bz)
    assert table[-1][1] == 28                                # the empty EOF block bgzip appends
    assert inflate.bgzf_block_table(gzip.compress(data)) is None
    assert inflate.bgzf_block_table(bz[:-5]) is None         # truncated chain is not trusted


def test_corrupt_streams_raise(tmp_path):
    data = _text(400_000, seed=6)
    bz = bytearray(inflate.bgzf_compress(data))
    bz[len(bz) // 2] ^= 0x55
    p = tmp_path / "bad.bgzf"
    p.write_bytes(bytes(bz))
    with pytest.raises(zlib.error):
        b"".join(inflate.iter_inflated(p, None, 2))
    p2 = tmp_path / "trunc.gz"
    p2.write_bytes(gzip.compress(data)[:-20])
    with pytest.raises(zlib.error):
        b"".join(inflate.iter_inflated(p2, None, 2))


def test_consumer_may_stop_early(tmp_path):
    p = tmp_path / "x.gz"
    p.write_bytes(gzip.compress(_text(2_000_000)))
    it = inflate.iter_inflated(p, None, 2)
    assert next(it).startswith(b"@r0\n")
    it.close()                                               # producer thread must not hang the interpreter


def test_pyfastx_stand_in_sniffs_like_the_reference_tests(tmp_path):
    D = ROOT / "tests" / "data"
    # /root/reference/tests/test_correct_workflow.py: fastq fixtures are fastq and not fasta, and vice versa
    assert fastx.is_fastq(D / "NC_001416.1_1.fq.gz") and not fastx.is_fasta(D / "NC_001416.1_1.fq.gz")
    assert fastx.is_fasta(D / "NC_001416.1.fasta.gz") and not fastx.is_fastq(D / "NC_001416.1.fasta.gz")
    (tmp_path / "lead.fa").write_bytes(b"\n\n  \n>r\nACGT\n")
    (tmp_path / "junk.txt").write_bytes(b"hello\n")
    (tmp_path / "empty").write_bytes(b"")
    assert fastx.is_fasta(tmp_path / "lead.fa")
    for f in ("junk.txt", "empty"):
        assert not fastx.is_fasta(tmp_path / f) and not fastx.is_fastq(tmp_path / f)
    with pytest.raises(RuntimeError):
        fastx.Fastq(tmp_path / "junk.txt", build_index=False)
    mod = fastx.install_pyfastx_stub()
    import pyfastx
    assert pyfastx is mod or hasattr(pyfastx, "Fasta")


@pytest.mark.parametrize("pad", [b"\0" * 512, b"\0", b"garbage after the last member\n" * 40])
def test_bytes_after_the_last_gzip_member_are_ignored_like_gzread(tmp_path, pad):
    """zlib's gzread (what mash reads its inputs with) and Python's gzip module both accept zero / tar padding after
    the last member; so does the host streamer (single member, several members, BGZF)."""
    data = _text(400_000, seed=5)
    for name, blob in (("single", gzip.compress(data, 3)),
                       ("multi", gzip.compress(data[:100_000], 1) + gzip.compress(data[100_000:], 1)),
                       ("bgzf", inflate.bgzf_compress(data))):
        p = tmp_path / name
        p.write_bytes(blob + pad)
        assert b"".join(inflate.iter_inflated(p, threads=2)) == data, name
        if pad[:1] == b"\0":
            assert gzip.decompress(blob + pad) == data       # the behaviour being matched


def test_bgzf_is_read_in_windows_not_as_one_blob(tmp_path, monkeypatch):
    """Memory stays flat: the BGZF path reads the file window by window (blocks straddle window boundaries), and a
    chain that stops being BGZF half way (here: an ordinary gzip member appended) continues on the sequential path."""
    data = _text(3_000_000, seed=6)
    blob = inflate.bgzf_compress(data)
    p = tmp_path / "w.gz"
    p.write_bytes(blob)
    monkeypatch.setattr(inflate, "BGZF_WINDOW", 100_003)
    reads = []
    real_open = open

    class Spy:
        def __init__(self, f):
            self.f = f

        def read(self, n=-1):
            reads.append(n)
            return self.f.read(n)

        def __getattr__(self, a):
            return getattr(self.f, a)

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return self.f.__exit__(*a)

    monkeypatch.setattr("builtins.open", lambda *a, **k: Spy(real_open(*a, **k)))
    st = StreamStats()
    assert b"".join(inflate.iter_inflated(p, st, threads=3)) == data
    assert max(reads) <= 100_003 and -1 not in reads and len(reads) > len(blob) // 100_003
    assert st.inflate_s > 0
    tail = _text(200_000, seed=7)
    p.write_bytes(blob + gzip.compress(tail, 1))
    assert b"".join(inflate.iter_inflated(p, threads=3)) == data + tail
