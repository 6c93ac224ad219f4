"""Drop-in bodies for the AuriClass methods that used to shell out to `mash`.

`MashGpuMethods` carries the re-pointed versions of

  BasicAuriclass.run_mash_dist        /root/reference/auriclass/classes.py:67-119
  BasicAuriclass.get_error_bounds     classes.py:285-324
  FastqAuriclass.sketch_fastq_query   classes.py:542-616
  FastaAuriclass.sketch_fasta_query   classes.py:663-723
  general.check_dependencies          general.py:181-205

with the same names, attributes touched (`stdout`, `estimated_genome_size`,
`mash_output`, `query_sketch_path`), log tags and exceptions.  The library
results are rendered back into Mash's text and pushed through the reference's
own parsing expressions, so 6-significant-digit rounding and pandas' int/float
inference are inherited, not re-implemented.

`plugin.install()` grafts these onto the unchanged `auriclass` package;
`FastqMashWorkflow` / `FastaMashWorkflow` are minimal stand-alone carriers of the
same methods for environments where the reference package is absent (the GPU
test box): they own exactly the attributes the hot path reads and writes, none
of the QC logic (out of scope).
"""
from __future__ import annotations

import logging
import math
from io import StringIO
from pathlib import Path
from typing import Any, Dict, Hashable, List, Optional

import pandas as pd

from . import _lib
from .mash_gpu import (NO_RECORDS_STDERR, GpuSketch, MashGPU, dist_stdout, fmt_g, sketch_fasta_stderr,
                       sketch_reads_stderr)
from .msh import read_msh

_ENGINES: Dict[int, MashGPU] = {}


def engine(device: int = 0) -> MashGPU:
    """Process-wide device context (AuriClass is single threaded: classes.py has no threads)."""
    e = _ENGINES.get(device)
    if e is None or not e._ctx:
        e = MashGPU(device)
        _ENGINES[device] = e
    return e


def process_group_world() -> int:
    """Ranks of the default torch.distributed process group when the caller set one up (one process per GPU, e.g. under
    torchrun), else 1.  torch is not imported for the ordinary single-process CLI."""
    import sys
    dist = sys.modules.get("torch.distributed")
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return 1
    return int(dist.get_world_size())


def add_tag(tag: str, lines: str) -> str:
    """Same rendering as auriclass/general.py:10-32."""
    full_tag = f"[{tag}]"
    if lines:
        return "\n".join([f"{full_tag} {line}" for line in lines.split("\n") if line != ""])
    return f"{full_tag}"


def check_dependencies() -> None:
    """Replaces the `mash -h` probe (general.py:198-205): the library and an sm_100 GPU must be there."""
    try:
        lib = _lib.load()
    except (ImportError, OSError, AttributeError) as e:
        raise FileNotFoundError(f"libauriclass_cuda.so is not installed: {e}")
    if lib.ac_probe() != 0:
        raise FileNotFoundError("libauriclass_cuda.so found but unusable: " + lib.ac_last_error().decode())


# ---------------------------------------------------------------------------
# `mash bounds -k K -p P` text (classes.py:305-318).  Host-side table of 2x9x8
# binomial quantiles: not on the GPU path.
# ---------------------------------------------------------------------------
def _binom_quantile_gt(n: int, p: float, q2: float) -> int:
    """Smallest x >= 0 with BinomialCDF(x; n, p) > q2, capped at n."""
    from scipy.stats import binom  # host-side utility only

    if p >= 1.0:
        return n
    x = int(binom.ppf(q2, n, p))  # smallest x with cdf >= q2 (up to rounding); settle the strict inequality
    x = max(x - 2, 0)
    while x < n and not (binom.cdf(x, n, p) > q2):
        x += 1
    return x


def mash_bounds_text(k: int, prob: float) -> str:
    sizes = [100, 500, 1000, 5000, 10000, 50000, 100000, 500000, 1000000]
    dists = [0.05, 0.1, 0.15, 0.2, 0.25, 0.3, 0.35, 0.4]
    q2 = (1.0 - prob) / 2.0
    out = [f"\nParameters (run with -h for details):\n   k:   {k}\n   p:   {fmt_g(prob)}\n\n"]
    for tbl in range(2):
        out.append("\t%s\n" % ("Mash distance" if tbl == 0 else "Screen distance"))
        out.append("Sketch" + "".join("\t" + fmt_g(d) for d in dists) + "\n")
        for s in sizes:
            row = [str(s)]
            for d in dists:
                m2j = 1.0 / (2.0 * math.exp(k * d) - 1.0) if tbl == 0 else (1.0 - d) ** k
                x = _binom_quantile_gt(s, m2j, q2)
                je = x / s
                if tbl == 0:
                    j2m = math.inf if je == 0 else -1.0 / k * math.log(2.0 * je / (1.0 + je))
                else:
                    j2m = 1.0 - je ** (1.0 / k)
                row.append(fmt_g(j2m - d))
            out.append("\t".join(row) + "\n")
        out.append("\n")
    return "".join(out)


# ---------------------------------------------------------------------------
class MashGpuMethods:
    """Method bodies; `self` is a BasicAuriclass-shaped object."""

    device: int = 0

    # -- mash sketch -r -m M -o Q -k K -s S files...   (classes.py:542-616)
    def sketch_fastq_query(self) -> None:
        command = ["mash", "sketch", "-r", "-m", str(self.minimal_kmer_coverage), "-o", self.query_sketch_path,
                   "-k", str(self.kmer_size), "-s", str(self.sketch_size), *self.read_paths]
        command_list_of_str = [str(item) for item in command]
        logging.info(add_tag("mash sketch", " ".join(command_list_of_str) + "   [libauriclass_cuda]"))
        world = process_group_world()
        if world > 1:
            # one process per GPU: read blocks are dealt over the ranks, the counts merged by one hash-range exchange
            # (BASELINE.json configs[2]); collective -- every rank runs this method and gets the same sketch
            import torch
            self._gpu_device = torch.cuda.current_device()
        eng = engine(getattr(self, "_gpu_device", self.device))
        try:
            if world > 1:
                sk = eng.sketch_reads_sharded([str(p) for p in self.read_paths], k=self.kmer_size, s=self.sketch_size,
                                              m=self.minimal_kmer_coverage)
            else:
                sk = eng.sketch_reads([str(p) for p in self.read_paths], k=self.kmer_size, s=self.sketch_size,
                                      m=self.minimal_kmer_coverage)
            stderr = sketch_reads_stderr(sk, self.query_sketch_path)
        except _lib.NoRecordsError:
            sk = None
            stderr = NO_RECORDS_STDERR.format("input files")
        if "ERROR: Did not find fasta records in" in stderr:
            raise ValueError(
                f"Did not find sequence records in {self.read_paths}. Please check if these are valid fastq files")
        self.stdout = ""
        self._gpu_query_sketches = [sk]
        self._gpu_query_key = str(self.query_sketch_path)
        if stderr:
            for line in stderr.splitlines():
                logging.info(add_tag("mash sketch", line))
        for line in stderr.splitlines():
            if "Estimated genome size" in line:
                self.estimated_genome_size = float(line.split()[-1])
                break
        else:
            raise ValueError("Estimated genome size could not be parsed from mash sketch STDERR")

    # -- mash sketch -o Q -k K -s S files...   (classes.py:663-723): one sketch per file
    def sketch_fasta_query(self) -> None:
        command = ["mash", "sketch", "-o", self.query_sketch_path, "-k", str(self.kmer_size), "-s",
                   str(self.sketch_size), *self.read_paths]
        command_list_of_str = [str(item) for item in command]
        logging.info(add_tag("mash sketch", " ".join(command_list_of_str) + "   [libauriclass_cuda]"))
        eng = engine(self.device)
        sketches: List[GpuSketch] = []
        try:
            for p in self.read_paths:
                sketches.append(eng.sketch_fasta(str(p), k=self.kmer_size, s=self.sketch_size))
            stderr = sketch_fasta_stderr([str(p) for p in self.read_paths], self.query_sketch_path)
        except _lib.NoRecordsError:
            stderr = NO_RECORDS_STDERR.format(str(self.read_paths[len(sketches)]))
        if "ERROR: Did not find fasta records in" in stderr:
            raise ValueError(
                f"Did not find sequence records in {self.read_paths}. Please check if these are valid fastq files")
        self.stdout = ""
        self._gpu_query_sketches = sketches
        self._gpu_query_key = str(self.query_sketch_path)
        if stderr:
            for line in stderr.splitlines():
                logging.info(add_tag("mash sketch", line))

    # -- FastaAuriclass.parse_genome_size (classes.py:725-751): sum of pyfastx.Fasta(path).size over the inputs.
    #    The GPU parser counted the bases while sketching, so this reads them back instead of re-indexing the file.
    def parse_genome_size(self) -> None:
        from . import fastx
        total_size = 0
        for filepath in self.read_paths:
            total_size += fastx.Fasta(filepath, build_index=True).size
        self.estimated_genome_size = total_size

    # -- mash dist REF Q   (classes.py:67-119)
    def run_mash_dist(self) -> pd.DataFrame:
        command = ["mash", "dist", self.reference_sketch_path, self.query_sketch_path]
        command_list_of_str = [str(item) for item in command]
        logging.info(add_tag("mash dist", " ".join(command_list_of_str) + "   [libauriclass_cuda]"))
        eng = engine(getattr(self, "_gpu_device", self.device))
        refs = eng.load_references(self.reference_sketch_path)
        queries = getattr(self, "_gpu_query_sketches", None)
        if not queries or getattr(self, "_gpu_query_key", None) != str(self.query_sketch_path):
            # a query .msh produced elsewhere: read it on the host and upload its hash lists
            q = read_msh(self.query_sketch_path)
            queries = [eng.sketch_from_hashes(r.hashes, q.kmer_size, q.sketch_size, r.length, r.name)
                       for r in q.references]
        stdout = dist_stdout(eng.dist(refs, queries))
        df = pd.read_csv(StringIO(stdout), sep="\t", header=None,
                         names=["Reference", "Query", "Distance", "P-value", "Matching-hashes"])
        df["Clade"] = df["Reference"].map(self.clade_dict["clade"])
        self.mash_output = df
        return df

    # -- mash bounds -k K -p P   (classes.py:285-324)
    def get_error_bounds(self) -> str:
        command = ["mash", "bounds", "-k", str(self.kmer_size), "-p", str(self.probability)]
        logging.info(add_tag("mash bounds", " ".join(command) + "   [auriclass_b200 host table]"))
        return mash_bounds_text(self.kmer_size, self.probability)


class _HotPathCarrier(MashGpuMethods):
    """Owns exactly the attributes of BasicAuriclass.__init__ (classes.py:16-65) the hot path uses."""

    def __init__(self, name: str, read_paths: List[Path], reference_sketch_path: Path, kmer_size: int = 27,
                 sketch_size: int = 50_000, minimal_kmer_coverage: int = 3,
                 clade_config_path: Optional[Path] = None, device: int = 0) -> None:
        self.name = name
        self.read_paths = read_paths
        self.reference_sketch_path = reference_sketch_path
        self.kmer_size = kmer_size
        self.sketch_size = sketch_size
        self.minimal_kmer_coverage = minimal_kmer_coverage
        self.probability = 0.99
        self.clade_dict: Dict[Hashable, Any] = (
            pd.read_csv(clade_config_path, index_col=0, dtype=str).to_dict(orient="dict")
            if clade_config_path else {"clade": {}})
        self.query_sketch_path: Path = Path()
        self.estimated_genome_size: float = float()
        self.stdout: str = ""
        self.stderr: str = ""
        self.mash_output: pd.DataFrame = pd.DataFrame()
        self.device = device


class FastqMashWorkflow(_HotPathCarrier):
    """sketch_fastq_query() + run_mash_dist(), as FastqAuriclass.run() calls them (classes.py:618-625)."""

    def run_hot_path(self) -> pd.DataFrame:
        self.query_sketch_path = Path("tmpfile.msh")
        self.sketch_fastq_query()
        return self.run_mash_dist()


class FastaMashWorkflow(_HotPathCarrier):
    """sketch_fasta_query() + run_mash_dist(), as FastaAuriclass.run() calls them (classes.py:753-760)."""

    def run_hot_path(self) -> pd.DataFrame:
        self.query_sketch_path = Path("tmpfile.msh")
        self.sketch_fasta_query()
        return self.run_mash_dist()
