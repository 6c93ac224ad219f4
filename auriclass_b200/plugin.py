"""Graft the B200 Mash path onto an unchanged `auriclass` installation.

    import auriclass_b200.plugin as p; p.install()

replaces, on the reference's own classes, the four methods that spawn `mash`
(/root/reference/auriclass/classes.py:67-119, :285-324, :542-616, :663-723) and
`general.check_dependencies` (general.py:181-205, also re-exported into
`auriclass.main`).  Everything else -- CLI, QC decisions, report.tsv writer,
logging -- keeps running the reference's code, so reports stay byte-identical.
See INTEGRATION.md for the one-line change a maintainer would commit instead.
"""
from __future__ import annotations

import importlib
import sys

from .workflow import MashGpuMethods, check_dependencies


def install(device: int = 0) -> None:
    from . import fastx
    fastx.install_pyfastx_stub()   # the unchanged package imports pyfastx at module level (classes.py:9, general.py:7)
    classes = importlib.import_module("auriclass.classes")
    general = importlib.import_module("auriclass.general")
    MashGpuMethods.device = device
    classes.BasicAuriclass.device = device
    classes.BasicAuriclass.run_mash_dist = MashGpuMethods.run_mash_dist
    classes.BasicAuriclass.get_error_bounds = MashGpuMethods.get_error_bounds
    classes.FastqAuriclass.sketch_fastq_query = MashGpuMethods.sketch_fastq_query
    classes.FastaAuriclass.sketch_fasta_query = MashGpuMethods.sketch_fasta_query
    classes.FastaAuriclass.parse_genome_size = MashGpuMethods.parse_genome_size
    general.check_dependencies = check_dependencies
    main = sys.modules.get("auriclass.main")
    if main is not None and hasattr(main, "check_dependencies"):
        main.check_dependencies = check_dependencies
