#!/usr/bin/env python
"""bench.py -- FASTQ Gbases/s sketched with the -m filter (BASELINE.json metric) on N B200s.

A step = one pass of the hot path (parse -> canonical k-mer MurmurHash3 -> exact min-copies
count -> bottom-s -> estimates -> dist against the 24 clade references) over one isolate's reads.

  --config c2 (default, the driver's line): BASELINE configs[1] -- a single R1 file of 150 bp synthetic reads at
      ~50x of a 12.4 Mb genome (4,133,334 reads, 620 Mbp, 1.33 GB of 4-line FASTQ), k=27 s=50000 m=3.  For N>1
      every rank sketches its own C2-sized shard of the same isolate (WEAK scaling) and the ranks merge by hash
      range (sharded.py).
  --config c3: BASELINE configs[2] -- R1 7.0 M + R2 7.0 M + unpaired 2.53 M reads (~200x, 2.48 Gbp, 5.4 GB of text,
      seeds SURVEY.md section 8d), FIXED total size, cut into read blocks, block b -> rank b mod N (STRONG scaling).

  value      : device-resident throughput (text already in HBM); CUDA events on the compute stream around each step
  e2e        : same metric through the C ABI from pinned HOST memory: H2D copies of the text, sketch, dist against
               the 24 clade references, D2H of the rows, wall clock
  parity     : the sketch of the exact timed input (device-fed and host-fed) against the CPU oracle over ALL ranks'
               reads -- hashes and counts bit for bit; a mismatch makes the run exit non-zero
  roofline   : dominant kernel (k2_hash): algorithmic bytes / CUDA-event time vs measured HBM peak;
               roofline_int: the binding roofline, 84 int32 instr per k-mer vs measured issue peak
  cpu_baseline / --impl reference : the oracle (CPU restatement of Mash) on the host cores

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

K, S, M = 27, 50000, 3
N_READS_C2 = 4_133_334
READ_LEN = 150
SEED_GENOME, SEED_READS = 20240001, 20240002
# C3: R1, R2, unpaired -- (reads, numpy PCG64 seed, file tag in the read names)
C3_FILES = ((7_000_000, 20240003, 1), (7_000_000, 20240004, 2), (2_533_334, 20240005, 3))
METRIC = "FASTQ Gbases/s sketched (-m filter)"
WORKLOADS = {   # the same string in both arms (the driver compares them); run specifics go to other config keys
    "c2": "C2: auriclass FASTQ mode, single R1 file, 150 bp synthetic reads at ~50x of a 12.4 Mb genome "
          "(4133334 reads, 620.0 Mbp, 1.331 GB 4-line FASTQ per GPU), k=27 s=50000 m=3 seed 42",
    "c3": "C3: auriclass FASTQ mode, R1 7.0 M + R2 7.0 M + unpaired 2.53 M 150 bp synthetic reads at ~200x of a 12.4 Mb "
          "genome (16533334 reads, 2480.0 Mbp, 5.37 GB 4-line FASTQ in total), read-block sharded, k=27 s=50000 m=3 seed 42",
}
I_ALG = 84  # int32 instructions per k-mer for MurmurHash3_x64_128 over 27 bytes, h1 only


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def bind_to_gpu_numa_node(local: int) -> str:
    """Pin this rank's threads to the CPUs of its GPU's NUMA node BEFORE the pinned staging memory is allocated
    (first touch puts the pages there), so H2D copies of N ranks do not all cross the socket interconnect."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        cpus = Path(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read_text().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return f"{bdf} -> cpus {cpus}"
    except Exception as e:  # no sysfs / attribute: leave the default placement
        return f"not bound ({type(e).__name__})"
    return "not bound"


# ----------------------------------------------------------------------------- inputs
def make_reads(n_reads: int, seed: int, tag: int, alloc):
    """C2 reads as 4-line FASTQ text, generated straight into a buffer from `alloc(nbytes)`."""
    from auriclass_b200 import synth
    contigs = synth.genome(SEED_GENOME)
    nbytes = n_reads * synth.record_bytes(READ_LEN, tag)
    buf = alloc(nbytes)
    synth.fastq_reads(contigs, n_reads, seed, READ_LEN, file_tag=tag, out=buf)
    return buf[:nbytes], contigs


def clade_names():
    p = ROOT / "data" / "clade_names.txt"
    return [l.strip() for l in p.read_text().splitlines() if l.strip()]


def reference_genomes(contigs):
    """24 references: 11 'auris' = G0 mutated at 0.05..1 %, 13 random outgroups."""
    from auriclass_b200 import synth
    out = []
    for i in range(11):
        rate = 0.0005 + (0.01 - 0.0005) * i / 10
        out.append(synth.mutate(contigs, rate, 20240010 + i))
    for i in range(13):
        rng = np.random.Generator(np.random.PCG64(20240030 + i))
        n = int(rng.integers(10_000_000, 16_000_000))
        out.append([np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n, dtype=np.uint8)]])
    return out


def build_reference_msh(path: Path, contigs, sketch_fn):
    from auriclass_b200 import synth
    from auriclass_b200.msh import MshReference, MshSketch, write_msh
    names = clade_names()
    refs = []
    for name, g in zip(names, reference_genomes(contigs)):
        h, length = sketch_fn(synth.fasta_bytes(g, prefix=name))
        refs.append(MshReference(name, "", int(length), h))
    write_msh(path, MshSketch(kmer_size=K, sketch_size=S, references=refs))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU baseline (oracle)
def cpu_baseline_sample(text: np.ndarray, n_reads_sample: int, rec: int, threads: int, n_reads_stream: int = 150_000):
    """The oracle (CPU restatement of Mash; the mash binary is not installed) on this box's host cores:
    all cores over `n_reads_sample` reads (value), plus the faithful single-thread streaming MinHashHeap
    -- what `mash sketch -r -m` without -p, i.e. AuriClass today, executes -- on a smaller sample.
    Returns (record, oracle sketch of the sample)."""
    from oracle import oracle
    sample = text[: n_reads_sample * rec]
    t0 = time.perf_counter()
    sk = oracle.sketch_parallel([sample], K, S, M, threads=threads)
    dt = time.perf_counter() - t0
    bases = n_reads_sample * READ_LEN
    out = {"value": bases / dt / 1e9, "unit": "Gbases/s", "cores": threads, "kind": "port",
           "sample": f"first {n_reads_sample} reads of the workload ({bases / 1e6:.1f} Mbp, {sample.size / 1e6:.0f} MB FASTQ), "
                     f"oracle.sketch_parallel (restated-Mash CPU baseline, not the mash binary), {dt:.2f} s wall = "
                     f"{dt * threads:.0f} core-seconds",
           "n_hashes": int(len(sk.hashes))}
    n1 = min(n_reads_sample, n_reads_stream)
    t0 = time.perf_counter()
    oracle.sketch_stream([text[: n1 * rec]], K, S, M)
    dt1 = time.perf_counter() - t0
    out["single_thread_stream"] = {"value": n1 * READ_LEN / dt1 / 1e9, "unit": "Gbases/s", "cores": 1,
                                   "sample": f"first {n1} reads, oracle.sketch_stream (faithful streaming MinHashHeap), {dt1:.2f} s"}
    return out, sk


def inflate_sample(text: np.ndarray, sample_bytes: int = 96 << 20) -> dict:
    """gzip inflate is the host's duty and is timed separately (north_star): inflate rate of this box on a sample of
    the workload as an ordinary single-member .gz (one zlib stream, what `mash` itself would do) and as BGZF
    (block-parallel on host threads, auriclass_b200/inflate.py), extrapolated to the whole C2 file."""
    import tempfile
    import zlib
    from auriclass_b200 import inflate
    from auriclass_b200.mash_gpu import StreamStats
    raw = text[:sample_bytes].tobytes()
    out = {"sample_bytes": len(raw), "extrapolated_to_bytes": int(text.size)}
    with tempfile.TemporaryDirectory() as d:
        for name, blob in (("single_member_gzip", zlib.compress(raw, 1, wbits=31)), ("bgzf", inflate.bgzf_compress(raw, level=1))):
            p = Path(d) / (name + ".gz")
            p.write_bytes(blob)
            st = StreamStats()
            t0 = time.perf_counter()
            n = sum(len(c) for c in inflate.iter_inflated(p, st))
            dt = time.perf_counter() - t0
            assert n == len(raw)
            out[name] = {"MBps": len(raw) / dt / 1e6, "threads": 1 if name.startswith("single") else inflate.default_threads(),
                         "wall_s_for_workload": text.size / (len(raw) / dt), "zlib_thread_seconds_sample": st.inflate_s}
    return out


def make_c3_files(alloc=lambda n: np.empty(n, np.uint8)):
    from auriclass_b200 import synth
    contigs = synth.genome(SEED_GENOME)
    files = []
    for n_reads, seed, tag in C3_FILES:
        buf = alloc(n_reads * synth.record_bytes(READ_LEN, tag))
        files.append(synth.fastq_reads(contigs, n_reads, seed, READ_LEN, file_tag=tag, out=buf))
    return files, contigs


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU implementation is entitled to every host core
    os.environ["OMP_NUM_THREADS"] = os.environ.get("AC_REF_THREADS", str(len(os.sched_getaffinity(0))))
    from auriclass_b200 import synth
    from oracle import oracle
    rec = synth.record_bytes(READ_LEN, 0)
    total_reads = args.reads if args.config == "c2" else sum(n for n, _, _ in C3_FILES)
    n_sample = args.ref_sample_reads or min(total_reads, N_READS_C2)
    if not args.ref_sample_reads and args.steps + args.warmup > 40:   # keep the whole run within a few minutes
        n_sample = max(100_000, n_sample * 40 // (args.steps + args.warmup))
    n_sample = min(total_reads, n_sample)
    text, contigs = make_reads(n_sample, SEED_READS if args.config == "c2" else C3_FILES[0][1], 0 if args.config == "c2" else 1,
                               lambda n: np.empty(n, np.uint8))
    cores = oracle.num_threads()
    bases = n_sample * READ_LEN
    # 24 reference sketches (dist is part of the per-isolate path); built once, outside the timed steps
    refs = []
    if not args.no_dist:
        for g in reference_genomes(contigs)[: args.n_refs]:
            sk = oracle.sketch_parallel([synth.fasta_bytes(g)], K, S, 1)
            refs.append((sk.hashes, sk.total_len))
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        sk = oracle.sketch_parallel([text], K, S, M)
        for h, length in refs:
            oracle.dist(h, sk.hashes, S, K, length, int(sk.genome_size))
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = bases / (ms * 1e-3) / 1e9
    desc = (f"{n_sample} reads x {READ_LEN} bp ({bases / 1e6:.1f} Mbp, {text.size / 1e6:.0f} MB FASTQ) per step: bounded sample of the "
            f"workload; oracle.sketch_parallel + {len(refs)} oracle.dist on {cores} host threads "
            "(restated-Mash CPU implementation; the mash binary is not available)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gbases/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak" if args.config == "c2" else "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "sample": desc},
        "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ----------------------------------------------------------------------------- our arm
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3"], help="c2: one C2 isolate per GPU (weak scaling, the driver's line); "
                    "c3: R1+R2+unpaired ~200x of fixed size, read-block sharded (strong scaling)")
    ap.add_argument("--reads", type=int, default=N_READS_C2, help="c2: reads per rank (default: the C2 workload)")
    ap.add_argument("--block-mb", type=int, default=32, help="c3: read-block size")
    ap.add_argument("--ref-sample-reads", type=int, default=0, help="reads per reference-arm step (0: one C2 isolate, shrunk if steps+warmup > 40)")
    ap.add_argument("--cpu-sample-reads", type=int, default=N_READS_C2)
    ap.add_argument("--n-refs", type=int, default=24)
    ap.add_argument("--no-dist", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the timed sketch (development runs only)")
    ap.add_argument("--feed-chunk-mb", type=int, default=32)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("note: timing rules ask for >= 3 warm-up steps")

    if args.impl == "reference":
        run_reference_arm(args)
        return

    # the contract is ONE JSON line on stdout: libraries that write to fd 1 (NCCL prints its version banner there)
    # are sent to stderr, the JSON line goes to a private duplicate of the original stdout
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: auriclass_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from auriclass_b200 import _lib, synth
    from auriclass_b200._lib import AC_MODE_READS, check
    from auriclass_b200.mash_gpu import MashGPU, StreamStats

    eng = MashGPU(local)
    lib = eng.lib
    rec = synth.record_bytes(READ_LEN, 0)

    # ---- inputs: this rank's reads in pinned host memory + one resident device copy
    t0 = time.perf_counter()
    full_files = None                      # c3, rank 0: the whole input (for the parity check)
    if args.config == "c2":
        nbytes = args.reads * rec
        pinned_ptr = lib.ac_alloc_pinned(nbytes)
        if not pinned_ptr:
            raise MemoryError("ac_alloc_pinned failed")
        host = np.ctypeslib.as_array((C.c_ubyte * nbytes).from_address(pinned_ptr))
        text, contigs = make_reads(args.reads, SEED_READS + 1000 * rank, 0, lambda n: host)
        reads_rank, reads_total = args.reads, args.reads * world
    else:
        from auriclass_b200.readblocks import blocks_of_rank, cut_read_blocks
        files, contigs = make_c3_files()
        blocks = cut_read_blocks(files, args.block_mb << 20)
        mine = blocks_of_rank(blocks, rank, world)
        nbytes = sum(b - a for _, a, b in mine)
        pinned_ptr = lib.ac_alloc_pinned(max(nbytes, 1))
        if not pinned_ptr:
            raise MemoryError("ac_alloc_pinned failed")
        host = np.ctypeslib.as_array((C.c_ubyte * max(nbytes, 1)).from_address(pinned_ptr))
        off = 0
        for fi, a, b in mine:               # every block starts at a record start and ends with a newline: the
            host[off:off + b - a] = files[fi][a:b]   # concatenation is 4-line FASTQ again
            off += b - a
        text = host[:nbytes]
        reads_rank = int(np.count_nonzero(text == 10)) // 4
        reads_total = sum(n for n, _, _ in C3_FILES)
        if rank == 0 and not args.no_parity:
            full_files = files
        else:
            del files
        log(f"[rank {rank}] c3: {len(mine)} of {len(blocks)} read blocks of <= {args.block_mb} MiB, {reads_rank} reads")
    dev_text = torch.empty(max(nbytes, 1), dtype=torch.uint8, device="cuda")
    dev_text[:nbytes].copy_(torch.from_numpy(text), non_blocking=False)
    torch.cuda.synchronize()
    log(f"[rank {rank}] numa: {numa}")
    log(f"[rank {rank}] generated {reads_rank} reads ({nbytes / 1e9:.3f} GB FASTQ) in {time.perf_counter() - t0:.1f} s")
    bases_rank, bases_total = reads_rank * READ_LEN, reads_total * READ_LEN

    # ---- 24 clade references: sketched by the library itself, written as .msh, read back by the host reader
    refs = None
    if not args.no_dist:
        t0 = time.perf_counter()
        tmp = Path(tempfile.mkdtemp()) / "synthetic_clade_references.msh"
        from auriclass_b200.msh import MshReference, MshSketch, write_msh
        rl = []
        for name, g in zip(clade_names()[: args.n_refs], reference_genomes(contigs)):
            sk = eng.sketch_bytes([synth.fasta_bytes(g, prefix=name)], k=K, s=S, m=1, mode=_lib.AC_MODE_FASTA)
            rl.append(MshReference(name, "", int(sk.length), sk.hashes()[0]))
            sk.free()
        write_msh(tmp, MshSketch(kmer_size=K, sketch_size=S, references=rl))
        refs = eng.load_references(tmp)
        log(f"[rank {rank}] built {len(rl)} reference sketches in {time.perf_counter() - t0:.1f} s")

    sharded = eng.sharder(None) if world > 1 else None
    timer_ms = C.c_float()

    def finish(job):
        if world > 1:
            return sharded.finish(job, K, S, M, name="reads_R1.fq")
        return eng._end(job, "reads_R1.fq", K, S, StreamStats())

    def step_device():
        """hot path with the text resident in HBM (what `value` times): sketch job + dist, device time by CUDA events"""
        check(lib.ac_timer_start(eng.ctx))
        job = eng._begin(K, S, 42, M, AC_MODE_READS)
        check(lib.ac_sketch_feed_device(job, dev_text.data_ptr(), nbytes, 0, 1))
        sk = finish(job)
        rows = eng.dist(refs, sk) if refs is not None else []
        check(lib.ac_timer_stop(eng.ctx, C.byref(timer_ms)))
        return sk, rows, float(timer_ms.value)

    chunk = args.feed_chunk_mb << 20

    def step_e2e():
        """C ABI from pinned host memory: H2D of the text every step, sketch, dist, rows back on the host"""
        job = eng._begin(K, S, 42, M, AC_MODE_READS)
        off = 0
        while True:
            n = min(chunk, nbytes - off)
            last = off + n >= nbytes
            check(lib.ac_sketch_feed(job, pinned_ptr + off, n, 0, 1 if last else 0))
            off += n
            if last:
                break
        sk = finish(job)
        rows = eng.dist(refs, sk) if refs is not None else []
        return sk, rows

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        check(lib.ac_sync(eng.ctx))

    # ---- warm-up
    for _ in range(args.warmup):
        sk, _, _ = step_device()
        sk.free()
    # ---- timed: device-resident
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = eng.launch_count()
    dev_ms, job_ms, t0 = [], [], time.perf_counter()
    dev_sketch = None
    for i in range(args.steps):
        sk, rows, ms = step_device()
        dev_ms.append(ms)
        job_ms.append(eng.last_job_times()["total_ms"])
        if i == args.steps - 1:
            dev_sketch = sk.hashes()
            dev_info = dict(sk.info)
        sk.free()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    launches = eng.launch_count() - l0
    # ---- timed: end to end from pinned host memory
    for _ in range(2):
        s_, _r = step_e2e()
        s_.free()
    barrier()
    t0 = time.perf_counter()
    rows, e2e_sketch = [], None
    for i in range(args.steps):
        s_, rows = step_e2e()
        if i == args.steps - 1:
            e2e_sketch = s_.hashes()
        s_.free()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    # the e2e path's own floor: one plain H2D copy of the same pinned bytes, all ranks copying AT THE SAME TIME (barrier
    # before every repetition), CUDA events, best of 3, slowest rank
    pinned_t = torch.from_numpy(text)
    h2d_ms = []
    for _ in range(3):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dev_text[:nbytes].copy_(pinned_t, non_blocking=True)
        e1.record()
        e1.synchronize()
        h2d_ms.append(e0.elapsed_time(e1))
    pcie_ms = min(h2d_ms)
    # dist alone (24 pairs): part of every step above, broken out for the stage table
    dist_ms = 0.0
    if refs is not None:
        sk, _, _ = step_device()
        check(lib.ac_timer_start(eng.ctx))
        eng.dist(refs, sk)
        check(lib.ac_timer_stop(eng.ctx, C.byref(timer_ms)))
        dist_ms = float(timer_ms.value)
        sk.free()

    # ---- kernel breakdown of one more (profiled) step; integer peak
    eng.set_profile(True)
    sk, _, _ = step_device()
    prof = eng.last_job_times()
    n_kmers, n_cand, passes = sk.info["n_kmers"], sk.info["n_candidates"], sk.info["passes"]
    if world > 1:   # the sharded sketch reports job-wide counters; the rooflines are per GPU
        n_kmers, n_cand = n_kmers // world, n_cand // world
    n_hashes = len(sk)
    sk.free()
    eng.set_profile(False)
    int_peak = eng.int_peak()
    pipe = eng.pipe_rates()

    # ---- max over ranks (device time of the step: every rank's own CUDA events)
    ms_step, ms_job = float(np.mean(dev_ms)), float(np.mean(job_ms))
    if world > 1:
        t = torch.tensor([ms_step, wall_ms, e2e_ms, pcie_ms, ms_job], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, wall_ms, e2e_ms, pcie_ms, ms_job = [float(x) for x in t.tolist()]

    # ---- parity: the sketch that was just timed against the CPU oracle over ALL ranks' reads
    parity = {"checked": False}
    cpu_rec = None
    if not args.no_parity:
        if args.config == "c2" and world > 1:     # rank 0 needs the other ranks' shards: over NVLink, then to the host
            gathered = [torch.empty_like(dev_text) for _ in range(world)] if rank == 0 else None
            dist.gather(dev_text, gathered, dst=0)
            if rank == 0:
                full_files = [g.cpu().numpy() for g in gathered]
                del gathered
        elif args.config == "c2":
            full_files = [text]
        if rank == 0:
            from oracle import oracle
            os.sched_setaffinity(0, all_cpus)   # the CPU oracle gets every host core again
            threads = len(all_cpus)
            t0 = time.perf_counter()
            if args.config == "c2" and world == 1 and not args.no_cpu_baseline and args.cpu_sample_reads >= args.reads:
                cpu_rec, osk = cpu_baseline_sample(text, args.reads, rec, threads)   # one oracle pass serves both
            else:
                osk = oracle.sketch_parallel(full_files, K, S, M, threads=threads)
            same = lambda got: bool(np.array_equal(got[0], osk.hashes) and np.array_equal(got[1], osk.counts))  # noqa: E731
            parity = {"checked": True, "n_hashes": int(len(osk.hashes)), "equal": same(dev_sketch) and same(e2e_sketch),
                      "device_fed_equal": same(dev_sketch), "host_fed_equal": same(e2e_sketch),
                      "genome_size_equal": bool(dev_info["genome_size"] == osk.genome_size),
                      "n_kmers_equal": bool(int(dev_info["n_kmers"]) == int(osk.n_kmers)),
                      "against": f"oracle.sketch_parallel over all {reads_total} reads of the {world} rank(s) "
                                 f"({threads} host threads, {time.perf_counter() - t0:.1f} s): hashes and counts bit for bit",
                      "sketch_xor": "%016x" % int(np.bitwise_xor.reduce(osk.hashes)) if len(osk.hashes) else "0"}
            parity["equal"] = bool(parity["equal"] and parity["genome_size_equal"] and parity["n_kmers_equal"])
            full_files = None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # K2 algorithmic bytes per step: the base stream (1 B per base + 1 separator per read) in, 8 B per candidate out
        k2_bytes = bases_rank + reads_rank + 8 * n_cand
        hash_s = max(prof["hash_ms"], 1e-6) * 1e-3
        # pipe-cycle roofline of K2: per k-mer MurmurHash3_x64_128 over 27 bytes needs 14 IMAD.WIDE +
        # 26 IMAD (12 64-bit multiplies by a constant = 1 wide + 2 low products each, two h*5+c) on the fmaheavy pipe and
        # 44 SHF/LOP3/IADD3 on the alu pipe; the two pipes run concurrently, each at its own measured rate
        fma_floor_s = n_kmers * (14.0 / pipe["imad_wide"] + 26.0 / pipe["imad"]) if min(pipe.values()) > 0 else 0.0
        alu_floor_s = n_kmers * 44.0 / pipe["lop3"] if pipe["lop3"] > 0 else 0.0
        floor_s = max(fma_floor_s, alu_floor_s)
        seg_mb = int(os.environ.get("AC_SEG_MB", "256"))
        n_seg = max(1, -(-nbytes // (seg_mb << 20)))  # segments per step (ac_common.cuh SEG_BYTES); K2 runs once per segment,
        k2_launches = n_seg + (1 if min(nbytes, seg_mb << 20) >= (32 << 20) else 0)   # and twice over the first one (ac_api.cu run_segment)
        traffic = None
        tp = ROOT / "profiles" / "k2_traffic.json"
        if tp.exists():
            try:
                tj = json.load(open(tp))
                traffic = tj.get("dram_bytes_per_launch")
                if traffic is not None:   # captured on a `segment_mb` segment: proportional to the bases per launch
                    traffic = int(traffic * (nbytes / n_seg) / (int(tj.get("segment_mb", seg_mb)) << 20))
            except Exception:
                traffic = None
        dev_timing = ("value: CUDA events (ac_timer_start / ac_timer_stop) on the compute stream around each whole step -- "
                      "sketch job, exchange and dist -- on every rank, mean over the steps, max over ranks")
        out = {
            "metric": METRIC,
            "value": bases_total / (ms_step * 1e-3) / 1e9 if args.config == "c3" else world * bases_rank / (ms_step * 1e-3) / 1e9,
            "unit": "Gbases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "wall_ms_per_step": wall_ms, "higher_is_better": True,
            "scaling": "weak" if args.config == "c2" else "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {
                "workload": WORKLOADS[args.config],
                "per_gpu": f"{reads_rank} reads, {bases_rank / 1e6:.1f} Mbp, {nbytes / 1e9:.3f} GB 4-line FASTQ on rank 0"
                           + (f"; read blocks of <= {args.block_mb} MiB dealt round-robin" if args.config == "c3" else ""),
                "l2": f"inputs larger than L2 ({nbytes / 1e9:.2f} GB text per GPU and step vs 126 MB)",
                "timing": dev_timing + "; wall_ms_per_step brackets the same steps with barrier + synchronize on both sides",
                "parallelism": "1 GPU" if world == 1 else f"read-block sharded over {world} GPUs, one hash-range exchange per isolate "
                                                          f"({sharded.transport})",
                "n_kmers_per_gpu": int(n_kmers), "n_candidates_per_gpu": int(n_cand), "sketch_hashes": int(n_hashes),
                "passes": int(passes),
            },
            "parity": parity,
            "e2e": {"value": (bases_total if args.config == "c3" else world * bases_rank) / (e2e_ms * 1e-3) / 1e9, "unit": "Gbases/s",
                    "ms_per_step": e2e_ms, "wall_s_per_isolate": e2e_ms * 1e-3,
                    "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(len(rows) * 32 + 8 * 256),
                    "pcie_floor_ms": pcie_ms, "pcie_GBps": nbytes / pcie_ms / 1e6, "frac_of_pcie_floor": pcie_ms / e2e_ms,
                    "pcie_floor": "one plain H2D copy of the same pinned bytes, all ranks copying concurrently (barrier first), slowest rank",
                    "includes": f"H2D of the FASTQ text from pinned memory in {args.feed_chunk_mb} MB feeds, sketch, dist vs "
                                f"{len(rows)} references, D2H of the rows; gzip inflate not included (host duty, see e2e.inflate)",
                    "stages_ms": {"h2d_floor": pcie_ms, "parse_k1": prof["parse_ms"], "hash_k2": prof["hash_ms"],
                                  "select_k3_k4": prof["select_ms"], "dist_k5_and_d2h": dist_ms,
                                  "exchange_and_host": max(0.0, ms_step - prof["parse_ms"] - prof["hash_ms"] - prof["select_ms"] - dist_ms)}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": k2_bytes / hash_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": k2_bytes / hash_s / 1e9 / hbm_peak, "traffic": traffic,
                         "kernel": "k2_hash<27, FLAT>", "launches_per_step": int(k2_launches), "segments_per_step": int(n_seg),
                         "avg_segment_ms": prof["hash_ms"] / n_seg, "algorithmic_bytes_per_segment": k2_bytes / n_seg,
                         "traffic_note": "DRAM bytes of one steady-state launch over a whole segment (ncu --set full, profiles/k2_traffic.json), "
                                         "scaled to this run's bytes per segment", "peak_source": peak_src,
                         "note": "K2 is integer-issue bound, not HBM bound: see roofline_int"},
            "roofline_int": {"bound": "int32-issue", "achieved": n_kmers * I_ALG / hash_s / 1e12, "peak": int_peak / 1e12,
                             "unit": "Tinstr/s", "frac": n_kmers * I_ALG / hash_s / int_peak,
                             "instr_per_kmer": I_ALG, "peak_source": "ac_int_peak: IMAD+LOP3 microbenchmark run in this process"},
            "roofline_pipe": {"bound": "fmaheavy pipe (IMAD / IMAD.WIDE)" if fma_floor_s >= alu_floor_s else "alu pipe (SHF/LOP3/IADD3)",
                              "floor_ms": floor_s * 1e3, "achieved_ms": hash_s * 1e3, "frac": floor_s / hash_s,
                              "fma_floor_ms": fma_floor_s * 1e3, "alu_floor_ms": alu_floor_s * 1e3,
                              "rates_Tinstr_per_s": {k: v / 1e12 for k, v in pipe.items()},
                              "model": "per k-mer 14 IMAD.WIDE + 26 IMAD (fmaheavy) and 44 SHF/LOP3/IADD3 (alu), pipes concurrent; "
                                       "rates measured in this process by ac_pipe_rates (one instruction kind issued alone)"},
            "sketch_only": {"value": (bases_total if args.config == "c3" else world * bases_rank) / (ms_job * 1e-3) / 1e9, "unit": "Gbases/s",
                            "ms_per_step": ms_job,
                            "note": "the sketch job alone (first to last kernel of the local pass, CUDA events, mean of the timed steps, max over "
                                    "ranks) -- what round 1 reported as `value`; `value` now also covers the exchange and the dist call"},
            "kernel_ms": {"parse_k1": prof["parse_ms"], "hash_k2": prof["hash_ms"], "select_k3_k4": prof["select_ms"],
                          "job_total": prof["total_ms"], "job_total_mean_timed": ms_job, "dist_k5_and_d2h": dist_ms},
        }
        if world == 1 and not args.no_cpu_baseline:   # (the contract times the CPU baseline at N = 1 only)
            os.sched_setaffinity(0, all_cpus)
            if cpu_rec is None:
                cpu_rec, _ = cpu_baseline_sample(text, min(reads_rank, args.cpu_sample_reads), rec, len(all_cpus))
            out["cpu_baseline"] = cpu_rec
            try:
                out["e2e"]["inflate"] = inflate_sample(text)
            except Exception as e:   # the inflate figure is an aside: never lose the line over it
                out["e2e"]["inflate"] = {"error": repr(e)}
        print(json.dumps(out), file=json_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    lib.ac_free_pinned(pinned_ptr)
    if rank == 0 and parity.get("checked") and not parity["equal"]:
        log("PARITY FAILURE: the timed sketch differs from the oracle: " + json.dumps(parity))
        sys.exit(3)


if __name__ == "__main__":
    main()
