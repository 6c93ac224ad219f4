"""Stand-in for the three things AuriClass asks of `pyfastx` on the Mash path:

* `is_fastq(file)` / `is_fasta(file)`  -- /root/reference/auriclass/general.py:68-115: does
  `pyfastx.Fastq(file, build_index=False)` / `pyfastx.Fasta(file, build_index=False)` accept the file?
  pyfastx decides on the first non-blank character of the (possibly gzip-compressed) stream: '@' / '>'.
* `Fasta(file, build_index=True).size` -- classes.py:746-751: total number of sequence characters of the
  assembly (genome-size QC of FASTA input).  That number is a by-product of the GPU record parser (kernel
  K1 counts every sequence byte it emits), so it costs no extra pass: `MashGPU.sketch_fasta` reports it in
  `GpuSketch.info["n_bases"]`; `Fasta.size` runs a sketch job only when nobody did so before.

`install_pyfastx_stub()` registers a module named `pyfastx` exposing exactly these names, so the unchanged
`auriclass` package imports in an environment that lacks the real one (as this image does).
"""
from __future__ import annotations

import gzip
import sys
import types
from pathlib import Path
from typing import Dict, Union

PathLike = Union[str, Path]
_sizes: Dict[str, int] = {}          # resolved path -> total bases, recorded by the sketch path


def first_record_char(path: PathLike) -> bytes:
    """First non-blank byte of the inflated stream (b"" for an empty file)."""
    with open(path, "rb") as f:
        gz = f.read(2) == b"\x1f\x8b"
    opener = gzip.open if gz else open
    with opener(path, "rb") as f:
        while True:
            chunk = f.read(4096)
            if not chunk:
                return b""
            stripped = chunk.lstrip(b" \t\r\n")
            if stripped:
                return stripped[:1]


def is_fastq(file: PathLike) -> bool:
    try:
        Fastq(file, build_index=False)
        return True
    except RuntimeError:
        return False


def is_fasta(file: PathLike) -> bool:
    try:
        Fasta(file, build_index=False)
        return True
    except RuntimeError:
        return False


def record_size(path: PathLike, n_bases: int) -> None:
    _sizes[str(Path(path).resolve())] = int(n_bases)


class Fastq:
    def __init__(self, file_name: PathLike, build_index: bool = True, **_):
        self.file_name = str(file_name)
        if first_record_char(file_name) != b"@":
            raise RuntimeError(f"{file_name} is not plain or gzip compressed fastq formatted file")


class Fasta:
    def __init__(self, file_name: PathLike, build_index: bool = True, **_):
        self.file_name = str(file_name)
        if first_record_char(file_name) != b">":
            raise RuntimeError(f"{file_name} is not plain or gzip compressed fasta formatted file")
        self._size = None

    @property
    def size(self) -> int:
        """Total sequence length of the file (all records)."""
        if self._size is None:
            key = str(Path(self.file_name).resolve())
            if key not in _sizes:
                from .workflow import engine
                sk = engine().sketch_fasta(self.file_name, k=21, s=1000)   # parser by-product; the sketch is discarded
                sk.free()
            self._size = _sizes[key]
        return self._size


def install_pyfastx_stub() -> types.ModuleType:
    """Make `import pyfastx` resolve to this stand-in when the real package is absent."""
    try:
        import pyfastx as real  # noqa: F401
        return real
    except ImportError:
        pass
    mod = types.ModuleType("pyfastx")
    mod.Fasta, mod.Fastq = Fasta, Fastq
    mod.__doc__ = "auriclass_b200.fastx stand-in for the pyfastx names AuriClass uses"
    sys.modules["pyfastx"] = mod
    return mod
