#!/usr/bin/env python
"""Phase timing of ShardedSketcher.finish on N GPUs (tuning aid): torchrun --nproc-per-node N tools/time_sharded.py"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from auriclass_b200 import synth  # noqa: E402
from auriclass_b200._lib import AC_MODE_READS, check  # noqa: E402
from auriclass_b200.mash_gpu import MashGPU  # noqa: E402
from auriclass_b200 import sharded as sh  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = MashGPU(local)
text = synth.fastq_reads(synth.genome(20240001), 4_133_334, 20240002 + 1000 * rank)
dev = torch.from_numpy(text).cuda()
S = sh.ShardedSketcher(eng, rank, world)
marks = []


def mark(name):
    torch.cuda.synchronize()
    marks.append((name, time.perf_counter()))


S._mark = mark
for it in range(5):
    marks.clear()
    dist.barrier(); torch.cuda.synchronize()
    mark("start")
    job = eng._begin(27, 50000, 42, 3, AC_MODE_READS)
    check(eng.lib.ac_sketch_feed_device(job, dev.data_ptr(), dev.numel(), 0, 1))
    mark("feed (launches queued)")
    sk = S.finish(job, 27, 50000, 3)
    mark("finish")
    sk.free()
    if rank == 0 and it >= 2:
        print(" | ".join(f"{n} {1e3 * (t - marks[i][1]):.2f}" for i, (n, t) in enumerate(marks[1:])), flush=True)
dist.barrier()
dist.destroy_process_group()
