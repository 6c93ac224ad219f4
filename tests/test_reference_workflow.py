"""The reference's own `run()` / CLI and its own test modules, UNCHANGED, with the B200 Mash path grafted on.

north_star: "the clade call and report.tsv must be identical".  The reference checks exactly that in CI with
`cmp report.tsv tests/data/reference_report_{fastq,fasta}.tsv` (/root/reference/.github/workflows/test.yaml:58-78)
and walks every QC attribute after the sketch and the dist in tests/test_correct_workflow.py:41-136, :139-234
(+ tests/test_failing_workflow.py:23-128 for missing / empty / non-FASTQ / mixed inputs).

`__graft_entry__.build()` stages the unmodified `auriclass` package and those two test modules under
oracle/_ref/ (git-ignored, shipped to the GPU box like the built .so); here they run in a subprocess with
`auriclass_b200.plugin.install()` applied first -- so everything but the four Mash methods is the reference's code:
argparse, validation, input-type guessing, QC decisions, the report writer.
"""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
STAGE = ROOT / "oracle" / "_ref"
DATA = ROOT / "tests" / "data"

needs_stage = pytest.mark.skipif(not (STAGE / "auriclass" / "main.py").exists(),
                                 reason="oracle/_ref not staged (run __graft_entry__.build() where /root/reference is mounted)")

CLI = """
import sys
sys.path.insert(0, {stage!r}); sys.path.insert(0, {root!r})
import auriclass_b200.plugin as p; p.install()
from auriclass.main import main
sys.argv = {argv!r}
main()
"""


def run_cli(argv, **kw):
    code = CLI.format(stage=str(STAGE), root=str(ROOT), argv=["auriclass", *argv])
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(ROOT), timeout=600, **kw)


@needs_stage
@pytest.mark.gpu
@pytest.mark.parametrize("kind,inputs", [
    ("fastq", ["tests/data/NC_001416.1_1.fq.gz", "tests/data/NC_001416.1_2.fq.gz"]),
    ("fasta", ["tests/data/NC_001416.1.fasta.gz"]),
])
def test_cli_report_tsv_is_byte_identical(kind, inputs, tmp_path):
    """Exactly the two CI invocations of test.yaml:58-78, then `cmp`."""
    report = tmp_path / "report.tsv"
    out = run_cli(["--expected_genome_size", "40000", "60000", "--output", str(report), "--log_file_path",
                   str(tmp_path / "run.log"), "--verbose", "-r", "tests/data/ref_sketch.msh", "-c", "tests/data/clade_config.csv", *inputs])
    assert out.returncode == 0, out.stderr[-3000:]
    golden = (DATA / f"reference_report_{kind}.tsv").read_bytes()
    assert report.read_bytes() == golden, (report.read_text(), golden.decode())
    # the library did the work: its tag is in the log next to Mash's command line
    log = (tmp_path / "run.log").read_text()
    assert "[libauriclass_cuda]" in log and "[mash dist]" in log
    if kind == "fastq":
        assert "[mash sketch] Estimated genome size: 48454.7" in log


@needs_stage
@pytest.mark.gpu
def test_reference_test_modules_pass_unchanged():
    """tests/test_correct_workflow.py and tests/test_failing_workflow.py of the reference, as they are."""
    for name in ("test_correct_workflow.py", "test_failing_workflow.py"):
        ref = Path("/root/reference/tests") / name
        if ref.exists():   # authoring container: the staged copy is the file itself
            assert ref.read_bytes() == (STAGE / "tests" / name).read_bytes()
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    out = subprocess.run([sys.executable, "-m", "pytest", str(STAGE / "tests"), "-q", "-p", "no:cacheprovider",
                          "--rootdir", str(STAGE), "-rA"], capture_output=True, text=True, cwd=str(ROOT), env=env, timeout=900)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    # 2 tests in test_correct_workflow.py + 4 in test_failing_workflow.py
    assert "6 passed" in out.stdout, tail


@needs_stage
@pytest.mark.gpu
def test_cli_multi_file_fasta_gives_one_block_of_rows_per_file(tmp_path):
    """`mash sketch` without -r makes one sketch per FASTA file; `mash dist` then prints n_refs rows per query file,
    query-major (classes.py:696-713, :92-118): 2 references x 2 files = 4 rows, and the unchanged run() digests them."""
    from auriclass_b200.workflow import FastaMashWorkflow
    wf = FastaMashWorkflow("multi", ["tests/data/NC_001416.1.fasta.gz", "tests/data/NC_001604.1.fasta.gz"],
                           "tests/data/ref_sketch.msh", clade_config_path="tests/data/clade_config.csv")
    df = wf.run_hot_path()
    assert list(df["Query"]) == ["tests/data/NC_001416.1.fasta.gz"] * 2 + ["tests/data/NC_001604.1.fasta.gz"] * 2
    assert list(df["Reference"]) == ["tests/data/NC_001416.1.fasta", "tests/data/NC_001604.1.fasta"] * 2
    assert list(df["Matching-hashes"]) == ["48476/48476", "0/50000", "0/50000", "39770/39770"]
    assert list(df["Distance"]) == [0, 1, 1, 0]
    report = tmp_path / "report.tsv"
    out = run_cli(["--expected_genome_size", "80000", "100000", "--output", str(report), "--log_file_path", str(tmp_path / "run.log"),
                   "-r", "tests/data/ref_sketch.msh", "-c", "tests/data/clade_config.csv",
                   "tests/data/NC_001416.1.fasta.gz", "tests/data/NC_001604.1.fasta.gz"])
    assert out.returncode == 0, out.stderr[-3000:]
    assert len(report.read_text().splitlines()) == 2


@needs_stage
def test_cli_without_a_gpu_stops_at_the_dependency_probe(tmp_path):
    """No CPU fallback: on a box without a GPU the unchanged CLI fails where it used to fail without `mash`
    (general.py:198-205 -> FileNotFoundError), before any sketching."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = run_cli(["--expected_genome_size", "40000", "60000", "--output", str(tmp_path / "r.tsv"), "--log_file_path",
                   str(tmp_path / "run.log"), "-r", "tests/data/ref_sketch.msh", "-c", "tests/data/clade_config.csv",
                   "tests/data/NC_001416.1.fasta.gz"])
    assert out.returncode != 0 and "FileNotFoundError" in out.stderr and "libauriclass_cuda.so" in out.stderr
    assert not (tmp_path / "r.tsv").exists()
    # --help / --version never needed mash, and still do not need the GPU
    assert run_cli(["--version"]).returncode == 0
