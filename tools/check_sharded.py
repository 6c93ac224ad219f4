#!/usr/bin/env python
"""Correctness and phase timing of the read-block sharded FASTQ path on N real GPUs (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 \
        tools/check_sharded.py [--out gpurun_out/check_sharded_nN.jsonl]

The driver's GPU test tier runs on one GPU, so what needs several ranks is checked here and the JSON lines are kept
under profiles/.  Every check compares with the CPU oracle (test infrastructure) over ALL reads on rank 0:

  fixture      MashGPU.sketch_reads_sharded on the reference's own read files (tests/data/NC_001416.1_{1,2}.fq.gz,
               /root/reference/tests/test_correct_workflow.py:20: both files -> one sketch), small read blocks so that
               every rank gets some, both transports (peer-memory kernels, torch.distributed collectives)
  workflow     the reference-shaped method (workflow.FastqMashWorkflow.sketch_fastq_query + run_mash_dist) under the
               process group: golden `mash dist` dict of the reference tests on every rank
  synthetic    three plain files (R1, R2, unpaired; ~100x of a 1.2 Mb genome) through the streaming cutter, count-table
               local passes, both transports
  failure      one rank's blocks are malformed: every rank raises, nobody hangs
  timing       exchange phases of a C2-sized shard per rank (device-resident), per transport
"""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from auriclass_b200 import _lib, synth  # noqa: E402
from auriclass_b200._lib import AC_MODE_READS, check  # noqa: E402
from auriclass_b200.mash_gpu import MashGPU  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--skip-timing", action="store_true")
ap.add_argument("--timing-reads", type=int, default=4_133_334)
args = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out_f = open(args.out, "a") if args.out and rank == 0 else None


def emit(rec):
    if rank == 0:
        line = json.dumps(rec)
        print(line, flush=True)
        if out_f:
            out_f.write(line + "\n")
            out_f.flush()


def all_true(flag: bool) -> bool:
    t = torch.tensor([1 if flag else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def same_on_all_ranks(h: np.ndarray, c: np.ndarray) -> bool:
    d = torch.tensor([int(np.bitwise_xor.reduce(h)) & 0x7FFFFFFFFFFFFFFF if len(h) else 0, int(c.astype(np.int64).sum()), len(h)],
                     device="cuda", dtype=torch.int64)
    lo, hi = d.clone(), d.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool((lo == hi).all().item())


def oracle_sketch(bufs, k, s, m):
    if rank != 0:
        return None
    from oracle import oracle
    return oracle.sketch_parallel(bufs, k, s, m, threads=len(os.sched_getaffinity(0)))


def engine_for(transport: str) -> MashGPU:
    os.environ["AC_XCHG"] = transport
    return MashGPU(local)


ok_all = True
data = ROOT / "tests" / "data"
fixture = [str(data / "NC_001416.1_1.fq.gz"), str(data / "NC_001416.1_2.fq.gz")]
tmp = Path(tempfile.gettempdir()) / "ac_check_sharded"
if rank == 0:
    tmp.mkdir(exist_ok=True)
    g = synth.genome(seed=20240001, scale=0.1)
    for i, (n, seed) in enumerate(((400_000, 31), (400_000, 32), (150_000, 33))):
        synth.fastq_reads(g, n, seed=seed, file_tag=i + 1).tofile(tmp / f"syn_{i}.fq")
    (tmp / "bad.fq").write_bytes(b"@r1\nACGTACGTACGTACGTACGTACGTACGTACGTACGT\nIIII\n" * 2000)
dist.barrier()
synthetic = [str(tmp / f"syn_{i}.fq") for i in range(3)]

for transport in ("peer", "collectives"):
    eng = engine_for(transport)
    sh = eng.sharder(None)
    # ---- fixture
    t0 = time.perf_counter()
    sk = eng.sketch_reads_sharded(fixture, k=27, s=50000, m=3, block_bytes=96 << 10)
    h, c = sk.hashes()
    from oracle import oracle as _o
    ref = oracle_sketch([_o.read_maybe_gz(p) for p in fixture], 27, 50000, 3) if rank == 0 else None
    good = same_on_all_ranks(h, c)
    if rank == 0:
        good = good and np.array_equal(h, ref.hashes) and np.array_equal(c, ref.counts) and sk.genome_size == ref.genome_size \
            and len(h) == 48451 and "%g" % sk.genome_size == "48454.7" and sk.info["n_kmers"] == ref.n_kmers
    good = all_true(good)
    ok_all &= good
    emit({"check": "fixture", "transport": sh.transport.split(":")[0].split("(")[0].strip(), "world": world, "ok": good,
          "n_hashes": len(h), "bytes_fed_rank0": sk.stats.bytes_in, "seconds": time.perf_counter() - t0})
    # ---- workflow method under the process group
    from auriclass_b200 import workflow
    workflow._ENGINES[local] = eng
    wf = workflow.FastqMashWorkflow("t", fixture, str(data / "ref_sketch.msh"), 27, 50_000, 3, str(data / "clade_config.csv"), device=local)
    wf.query_sketch_path = Path("tmpfile.msh")
    wf.sketch_fastq_query()
    df = wf.run_mash_dist()
    golden = json.load(open(ROOT / "tests" / "golden" / "workflow_dicts.json"))["mash_output_to_dict_fastq"]
    got = {k: {str(i): v for i, v in col.items()} for k, col in df.to_dict().items()}
    got["Query"] = {i: v.replace(str(ROOT) + "/", "") for i, v in got["Query"].items()}
    good = all_true(got == golden and wf.estimated_genome_size == 48454.7 and workflow.process_group_world() == world)
    ok_all &= good
    emit({"check": "workflow", "transport": transport, "world": world, "ok": good, "estimated_genome_size": wf.estimated_genome_size})
    # ---- synthetic, count-table local passes
    t0 = time.perf_counter()
    sk = eng.sketch_reads_sharded(synthetic, k=27, s=50000, m=3, block_bytes=8 << 20)
    h, c = sk.hashes()
    ref = oracle_sketch([np.fromfile(p, np.uint8) for p in synthetic], 27, 50000, 3)
    good = same_on_all_ranks(h, c)
    if rank == 0:
        good = good and np.array_equal(h, ref.hashes) and np.array_equal(c, ref.counts) and sk.genome_size == ref.genome_size \
            and sk.coverage == ref.coverage and sk.info["n_kmers"] == ref.n_kmers and sk.info["n_records"] == ref.n_records
    good = all_true(good)
    ok_all &= good
    emit({"check": "synthetic", "transport": transport, "world": world, "ok": good, "n_hashes": len(h), "tau_cut": sh.last_exchange.get("tau_cut"),
          "received_rank0": sh.last_exchange.get("received"), "bytes_fed_rank0": sk.stats.bytes_in, "seconds": time.perf_counter() - t0})
    # ---- a rank with malformed blocks
    t0 = time.perf_counter()
    raised = ""
    try:
        eng.sketch_reads_sharded([synthetic[0], str(tmp / "bad.fq")], k=27, s=1000, m=1, block_bytes=32 << 10)
    except _lib.AcError as e:
        raised = type(e).__name__ + ":" + str(e.code)
    good = all_true(raised != "")
    ok_all &= good
    emit({"check": "failure", "transport": transport, "world": world, "ok": good, "rank0_raised": raised, "seconds": time.perf_counter() - t0})
    # ---- phases of a C2-sized shard per rank
    if not args.skip_timing:
        text = synth.fastq_reads(synth.genome(20240001), args.timing_reads, 20240002 + 1000 * rank)
        dev = torch.from_numpy(text).cuda()
        marks = []

        def mark(name):
            torch.cuda.synchronize()
            check(eng.lib.ac_sync(eng.ctx))
            marks.append((name, time.perf_counter()))

        sh._mark = mark
        rows = []
        for it in range(6):
            marks.clear()
            dist.barrier()
            mark("start")
            job = eng._begin(27, 50000, 42, 3, AC_MODE_READS)
            check(eng.lib.ac_sketch_feed_device(job, dev.data_ptr(), dev.numel(), 0, 1))
            mark("local_pass")
            sk = sh.finish(job, 27, 50000, 3)
            mark("finish")
            sk.free()
            if it >= 2:
                rows.append({n: 1e3 * (t - marks[i][1]) for i, (n, t) in enumerate(marks[1:])})
        sh._mark = None
        mean = {k: float(np.mean([r[k] for r in rows])) for k in rows[0]}
        tail = sum(v for k, v in mean.items() if k not in ("local_pass",))
        t = torch.tensor([tail], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        emit({"check": "timing", "transport": transport, "world": world, "phases_ms_rank0": mean,
              "exchange_tail_ms_max_over_ranks": float(t.item()),
              "note": "every mark synchronises the device, so the phases are serialised here (bench.py does not do that)"})
        del dev, text
    eng.close()

emit({"check": "summary", "world": world, "ok": bool(ok_all)})
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
