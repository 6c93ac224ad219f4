"""CPU tests of everything that is not a CUDA kernel: the device hash core compiled for the host,
the C-ABI surface, the .msh reader/writer, Mash text rendering and the bounds table."""
import ctypes as C
import json
import re
import sys
from io import StringIO
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from auriclass_b200 import _lib, msh
from auriclass_b200.mash_gpu import DistRow, dist_stdout, fmt_g
from auriclass_b200.workflow import mash_bounds_text
from oracle import oracle

ROOT = Path(__file__).resolve().parent.parent
D = "tests/data/"


# ----------------------------------------------------------------- kmer_core.cuh on the CPU
@pytest.fixture(scope="module")
def hostcheck():
    so = ROOT / "tests" / "hostcheck" / "libac_hostcheck.so"
    if not so.exists():
        import __graft_entry__ as g
        g.build()
    L = C.CDLL(str(so))
    L.hc_hash_stream.restype = C.c_long
    L.hc_hash_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
    L.hc_hash_stream_quad.restype = C.c_long
    L.hc_hash_stream_quad.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
    L.hc_sanitise.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    return L


def _core_hashes(L, seq: bytes, k: int, quad: bool = False) -> np.ndarray:
    a = np.frombuffer(seq, np.uint8).copy()
    z = np.zeros_like(a)
    L.hc_sanitise(a.ctypes.data, a.size, z.ctypes.data)
    out = np.zeros(max(1, a.size), np.uint64)
    fn = L.hc_hash_stream_quad if quad else L.hc_hash_stream
    n = fn(z.ctypes.data, z.size, k, out.ctypes.data, out.size)
    out = out[:n]
    return out if k > 16 else out & np.uint64(0xFFFFFFFF)


@pytest.mark.parametrize("k", [1, 3, 4, 5, 8, 11, 15, 16, 17, 21, 24, 27, 28, 31, 32])
def test_k2_flat_loop_keeps_every_hash_below_tau(hostcheck, k):
    """The flat hot loop (strand from the first four bases, flag on a tie or a passing gate, exact re-hash of the
    flagged windows) must deliver exactly the hashes <= tau, for thresholds from 1/32 of the space down to 0."""
    hostcheck.hc_hash_stream_flat.restype = C.c_long
    hostcheck.hc_hash_stream_flat.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p]
    rng = np.random.default_rng(100 + k)
    seq = bytearray(rng.choice(np.frombuffer(b"ACGT", np.uint8), 20_000).tobytes())
    # palindromic stretches (first four bases of both strands tie), Ns, a low-complexity run
    seq[500:540] = b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"
    seq[3000:3003] = b"NNN"
    seq[7000:7100] = b"AT" * 50
    seq[12000:12064] = b"AATT" * 16
    seq = bytes(seq)
    a = np.frombuffer(seq, np.uint8).copy()
    z = np.zeros_like(a)
    hostcheck.hc_sanitise(a.ctypes.data, a.size, z.ctypes.data)
    want_all = oracle.all_hashes(b">x\n" + seq + b"\n", k)
    space = 64 if k > 16 else 32
    for tau in ((1 << space) // 32, (1 << space) // 250, (1 << space) - 1, 0):
        out = np.zeros(a.size, np.uint64)
        n_flagged = C.c_long(0)
        n = hostcheck.hc_hash_stream_flat(z.ctypes.data, z.size, k, tau, out.ctypes.data, out.size, C.byref(n_flagged))
        want = want_all[want_all <= np.uint64(tau)]
        assert n == want.size, (k, tau)
        assert np.array_equal(out[:n], want), (k, tau)
        if k >= 12 and tau == (1 << space) // 250:   # the filter is tight: candidates + first-word ties + a few
            assert n_flagged.value <= n + a.size // 256 * 2 + 400


@pytest.mark.parametrize("k", list(range(1, 33)))
def test_device_hash_core_matches_oracle_all_k(hostcheck, k):
    fa = oracle.read_maybe_gz(D + "NC_001416.1.fasta.gz")
    seq = b"".join(fa.split(b"\n")[1:])[:6000]
    assert np.array_equal(_core_hashes(hostcheck, seq, k), oracle.all_hashes(b">x\n" + seq + b"\n", k))
    # the quad layout kernel K2 uses (pre-shifted words, windows at a stride of 4 bytes per lane)
    assert np.array_equal(_core_hashes(hostcheck, seq[:3001], k, quad=True), oracle.all_hashes(b">x\n" + seq[:3001] + b"\n", k))


def test_device_hash_core_invalid_bases(hostcheck):
    rng = np.random.default_rng(7)
    s = rng.choice(list(b"ACGTacgtNnRY-*"), size=20000,
                   p=[.2, .2, .2, .2, .03, .03, .03, .03, .02, .01, .01, .01, .015, .015]).astype(np.uint8).tobytes()
    for k in (27, 17, 16, 9, 32):
        assert np.array_equal(_core_hashes(hostcheck, s, k), oracle.all_hashes(b">x\n" + s + b"\n", k))
        assert np.array_equal(_core_hashes(hostcheck, s, k, quad=True), oracle.all_hashes(b">x\n" + s + b"\n", k))


def test_device_hash_core_full_genome_k27(hostcheck):
    fa = oracle.read_maybe_gz(D + "NC_001416.1.fasta.gz")
    seq = b"".join(fa.split(b"\n")[1:])
    got = np.sort(_core_hashes(hostcheck, seq, 27))
    assert np.array_equal(got, msh.read_msh(D + "ref_sketch.msh").references[0].hashes)


# ----------------------------------------------------------------- C ABI surface
def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "auriclass_cuda.h").read_text()
    declared = set(re.findall(r"\b(ac_[a-z0-9_]+)\s*\(", header))
    declared -= {"ac_ctx", "ac_job", "ac_sketch", "ac_refs"}
    lib = _lib.load()  # binds every name in SIGNATURES; AttributeError if one is missing
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.ac_abi_version() == 1


def test_sketch_info_struct_layout():
    assert C.sizeof(_lib.SketchInfo) == 9 * 8 + 2 * 8 + 2 * 4 + 8


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    ctx = C.c_void_p()
    assert lib.ac_create(0, C.byref(ctx)) == _lib.AC_E_CUDA
    assert b"no CPU path" in lib.ac_last_error() or b"CUDA" in lib.ac_last_error()
    from auriclass_b200.workflow import check_dependencies
    with pytest.raises(FileNotFoundError):
        check_dependencies()


def test_plugin_grafts_onto_the_unchanged_reference_package():
    """auriclass_b200.plugin.install() on the reference package itself (only where /root/reference is mounted):
    the four subprocess methods, check_dependencies and parse_genome_size are re-pointed, and the package
    imports although pyfastx is absent (stand-in module)."""
    ref = Path("/root/reference")
    if not (ref / "auriclass" / "classes.py").exists():
        pytest.skip("reference package not mounted")
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import auriclass_b200.plugin as p; p.install()\n"
        "import auriclass.classes as c, auriclass.general as g\n"
        "from auriclass_b200.workflow import MashGpuMethods as M, check_dependencies as cd\n"
        "assert c.BasicAuriclass.run_mash_dist is M.run_mash_dist and c.BasicAuriclass.get_error_bounds is M.get_error_bounds\n"
        "assert c.FastqAuriclass.sketch_fastq_query is M.sketch_fastq_query and c.FastaAuriclass.sketch_fasta_query is M.sketch_fasta_query\n"
        "assert c.FastaAuriclass.parse_genome_size is M.parse_genome_size and g.check_dependencies is cd\n"
        "assert g.guess_input_type(['tests/data/NC_001416.1.fasta.gz']) == 'fasta'\n"
        "assert g.guess_input_type(['tests/data/NC_001416.1_1.fq.gz', 'tests/data/NC_001416.1_2.fq.gz']) == 'fastq'\n"
        "print('ok')\n") % (str(ref), str(ROOT))
    import subprocess
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(ROOT))
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr[-2000:]


def test_product_never_imports_oracle():
    for p in (ROOT / "auriclass_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".h"):
            txt = p.read_text()
            assert "import oracle" not in txt and "from oracle" not in txt and "libmash_oracle" not in txt, p


# ----------------------------------------------------------------- .msh reader / writer
def test_msh_reader_fixture_checksums():
    m = msh.read_msh(D + "ref_sketch.msh")
    assert (m.kmer_size, m.sketch_size, m.seed, m.hash_bits, m.alphabet) == (27, 50000, 42, 64, "ACGT")
    lam, t7 = m.references
    assert (lam.name, lam.length, len(lam.hashes)) == ("tests/data/NC_001416.1.fasta", 48502, 48476)
    assert (t7.name, t7.length, len(t7.hashes)) == ("tests/data/NC_001604.1.fasta", 39937, 39770)
    # SURVEY.md App. B checksums
    assert int(lam.hashes.min()) == 0x000231CD9AB4EFF9 and int(lam.hashes.max()) == 0xFFFB03E04778B43C
    assert int(np.bitwise_xor.reduce(lam.hashes)) == 0x5342B54818A13A47
    assert int(lam.hashes.sum(dtype=np.uint64)) == 0x50103BB97AA43FB5
    assert int(np.bitwise_xor.reduce(t7.hashes)) == 0x619A210ABF2D07C4
    assert np.all(np.diff(lam.hashes.astype(np.float64)) > 0)


def test_msh_writer_roundtrip(tmp_path):
    m = msh.read_msh(D + "ref_sketch.msh")
    m.references[0].counts = np.arange(len(m.references[0].hashes), dtype=np.uint32)
    p = tmp_path / "rt.msh"
    msh.write_msh(p, m)
    r = msh.read_msh(p)
    assert (r.kmer_size, r.sketch_size, r.seed, r.hash_bits) == (27, 50000, 42, 64)
    for a, b in zip(m.references, r.references):
        assert a.name == b.name and a.length == b.length and np.array_equal(a.hashes, b.hashes)
    assert np.array_equal(r.references[0].counts, m.references[0].counts)
    m32 = msh.MshSketch(kmer_size=16, sketch_size=100, hash_bits=32, seed=7,
                        references=[msh.MshReference("a", "c", 5000, np.arange(1, 101, dtype=np.uint64) * 1000)])
    msh.write_msh(p, m32)
    r = msh.read_msh(p)
    assert (r.kmer_size, r.seed, r.hash_bits) == (16, 7, 32) and np.array_equal(r.references[0].hashes, m32.references[0].hashes)


def test_msh_rejects_garbage(tmp_path):
    p = tmp_path / "x.msh"
    p.write_bytes(b"\x00\x01")
    with pytest.raises(ValueError):
        msh.read_msh(p)


# ----------------------------------------------------------------- Mash text
def test_fmt_g_matches_mash_printing():
    assert fmt_g(9.554047e-06) == "9.55405e-06" and fmt_g(48454.7123) == "48454.7"
    assert fmt_g(0.0) == "0" and fmt_g(1.0) == "1" and fmt_g(154153.2) == "154153" and fmt_g(1e-300) == "1e-300"


def test_dist_stdout_feeds_reference_parser_with_same_dtypes():
    # pandas infers int64 when every distance prints as an integer: report prints `0`, not `0.0`
    golden = json.load(open("tests/golden/workflow_dicts.json"))
    rows = [DistRow("tests/data/NC_001416.1.fasta", "tests/data/NC_001416.1.fasta.gz", 0.0, 0.0, 48476, 48476),
            DistRow("tests/data/NC_001604.1.fasta", "tests/data/NC_001416.1.fasta.gz", 1.0, 1.0, 0, 50000)]
    df = pd.read_csv(StringIO(dist_stdout(rows)), sep="\t", header=None,
                     names=["Reference", "Query", "Distance", "P-value", "Matching-hashes"])
    assert df["Distance"].dtype == np.int64
    clade = pd.read_csv(D + "clade_config.csv", index_col=0, dtype=str).to_dict(orient="dict")
    df["Clade"] = df["Reference"].map(clade["clade"])
    want = {k: {int(i): v for i, v in col.items()} for k, col in golden["mash_output_to_dict_fasta"].items()}
    assert df.to_dict() == want
    rows = [DistRow("tests/data/NC_001416.1.fasta", "tests/data/NC_001416.1_1.fq.gz", 9.554047e-06, 0.0, 48451, 48476),
            DistRow("tests/data/NC_001604.1.fasta", "tests/data/NC_001416.1_1.fq.gz", 1.0, 1.0, 0, 50000)]
    df = pd.read_csv(StringIO(dist_stdout(rows)), sep="\t", header=None,
                     names=["Reference", "Query", "Distance", "P-value", "Matching-hashes"])
    df["Clade"] = df["Reference"].map(clade["clade"])
    want = {k: {int(i): v for i, v in col.items()} for k, col in golden["mash_output_to_dict_fastq"].items()}
    assert df.to_dict() == want


def test_bounds_text_matches_reference_golden():
    # `mash bounds -k 27 -p 0.99` text pinned at /root/reference/tests/test_correct_workflow.py:26
    golden = open("tests/golden/mash_bounds_k27_p0.99.txt").read()
    assert mash_bounds_text(27, 0.99) == golden
    assert mash_bounds_text(21, 0.95) == oracle.bounds_text(21, 0.95)
