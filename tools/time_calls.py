#!/usr/bin/env python
"""Wall-clock of every C-ABI call of one device-resident sketch job (tuning aid; not part of the bench contract)."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from auriclass_b200 import _lib, synth  # noqa: E402
from auriclass_b200._lib import AC_MODE_READS, check  # noqa: E402
from auriclass_b200.mash_gpu import MashGPU, StreamStats  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_133_334
eng = MashGPU(0)
lib = eng.lib
contigs = synth.genome(20240001)
text = synth.fastq_reads(contigs, n_reads, 20240002)
dev = torch.from_numpy(text).cuda()
torch.cuda.synchronize()
for it in range(6):
    t = [time.perf_counter()]
    job = eng._begin(27, 50000, 42, 3, AC_MODE_READS); t.append(time.perf_counter())
    check(lib.ac_sketch_feed_device(job, dev.data_ptr(), dev.numel(), 0, 1)); t.append(time.perf_counter())
    sk = eng._end(job, "x", 27, 50000, StreamStats()); t.append(time.perf_counter())
    sk.free(); t.append(time.perf_counter())
    ms = (C.c_float * 4)()
    lib.ac_last_job_times(eng.ctx, ms)
    d = [1e3 * (b - a) for a, b in zip(t, t[1:])]
    print(f"begin {d[0]:.2f} feed {d[1]:.2f} end {d[2]:.2f} free {d[3]:.2f} ms | device job {ms[3]:.2f} ms", flush=True)
