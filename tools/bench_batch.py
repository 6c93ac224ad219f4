#!/usr/bin/env python
"""Throughput AND correctness of the two batch paths (BASELINE.json configs[3] and configs[4]) on N GPUs -- secondary
measurements beside bench.py's headline metric.  One JSON line per path (rank 0).

    python tools/bench_batch.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 \
        tools/bench_batch.py --assemblies 1000 --sketches 8192 --out gpurun_out/batch_n8.jsonl

  c4: batch classification -- N synthetic 12.4 Mb assemblies (G0 mutated at U(0, 2 %), seeds 20240100 + i, SURVEY.md
      section 8d), query q -> rank q mod P, each sketched from HOST memory (H2D inside the timed region) and distanced
      against 24 references; assemblies/s over all ranks (wall clock between barriers).  A sample of the queries is
      re-done by the CPU oracle on rank 0: sketch hashes, shared-hash counts and denominators bit for bit, distances
      to 1e-12.
  c5: all-vs-all dist -- T x T tiles of the upper triangle of synthetic 50,000-hash sketches (clusters sharing a base
      set, section 8d), tiles dealt round-robin over the ranks, the sketch set replicated on every GPU; pairs/s over
      all ranks, extrapolated to the 50,000 x 50,000 problem (1.25e9 unordered pairs).  A sample of the pairs is
      checked against the oracle's sequential capped merge.
"""
import argparse
import ctypes
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from auriclass_b200 import synth  # noqa: E402
from auriclass_b200._lib import AC_MODE_FASTA  # noqa: E402
from auriclass_b200.batch import all_pairs, classify_batch, shard_indices, tile_plan  # noqa: E402
from auriclass_b200.mash_gpu import MashGPU  # noqa: E402

K, S = 27, 50000


def c5_sketches(n, s, seed=20240500, clusters=None):
    """n sketches in clusters that share a base set of 1.2 s hashes; every sketch replaces each element with probability
    p ~ U(0, 0.3) by a fresh value, sorts and keeps the s smallest (scaled: 500 clusters at n = 50,000)."""
    clusters = clusters or max(1, n // 100)
    rng = np.random.Generator(np.random.PCG64(seed))
    bases = [np.unique(rng.integers(0, 1 << 63, int(1.2 * s), dtype=np.uint64)) for _ in range(clusters)]
    out = []
    for i in range(n):
        h = bases[i % clusters].copy()
        rep = rng.random(h.size) < rng.uniform(0, 0.3)
        h[rep] = rng.integers(0, 1 << 63, int(rep.sum()), dtype=np.uint64)
        h.sort()
        h = h[np.concatenate(([True], h[1:] != h[:-1]))]
        out.append(h[:s])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--assemblies", type=int, default=24)
    ap.add_argument("--sketches", type=int, default=768)
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--skip", default="")
    ap.add_argument("--check", type=int, default=6, help="queries (c4) / 16 x pairs (c5) re-done by the CPU oracle")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()

    def emit(rec):
        if rank == 0:
            line = json.dumps(rec)
            print(line, flush=True)
            if args.out:
                with open(args.out, "a") as f:
                    f.write(line + "\n")

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng = MashGPU(local)

    class Stopwatch:
        """wall time per call of an engine method, and -- with device=True -- the CUDA-event time of what the call put
        on the compute stream (kernel + result copies), so that host overhead and kernel time can be told apart"""

        def __init__(self, name, device=False):
            self.name, self.device, self.wall_ms, self.dev_ms = name, device, [], []
            self.orig = getattr(eng, name)
            setattr(eng, name, self)

        def __call__(self, *a, **kw):
            if self.device:
                eng.lib.ac_timer_start(eng.ctx)
            t = time.perf_counter()
            out = self.orig(*a, **kw)
            self.wall_ms.append(1e3 * (time.perf_counter() - t))
            if self.device:
                ms = ctypes.c_float()
                eng.lib.ac_timer_stop(eng.ctx, ctypes.byref(ms))
                self.dev_ms.append(ms.value)
            return out

        def reset(self):
            self.wall_ms, self.dev_ms = [], []

        def stats(self):
            st = {"calls": len(self.wall_ms), "wall_ms_total": float(np.sum(self.wall_ms)), "wall_ms_median": float(np.median(self.wall_ms)),
                  "wall_ms_max": float(np.max(self.wall_ms))} if self.wall_ms else {"calls": 0}
            if self.dev_ms:
                st.update(device_ms_total=float(np.sum(self.dev_ms)), device_ms_median=float(np.median(self.dev_ms)))
            return st

        def restore(self):
            setattr(eng, self.name, self.orig)
    if "c4" not in args.skip:
        g0 = synth.genome(20240001)
        refs_sk = []
        for i in range(24):
            fa = synth.fasta_bytes(synth.mutate(g0, 0.0005 + 0.0004 * i, 20240010 + i))
            sk = eng.sketch_bytes([fa], k=K, s=S, m=1, mode=AC_MODE_FASTA)
            refs_sk.append((sk.hashes()[0], sk.length))
            sk.free()
        refs = eng.load_sketch_set([h for h, _ in refs_sk], [l for _, l in refs_sk], K, S, names=[f"ref{i}" for i in range(24)])
        n = args.assemblies
        mine = set(shard_indices(n, rank, world))

        def assembly(i):
            rate = float(np.random.Generator(np.random.PCG64(20240100 + i)).uniform(0, 0.02))
            return synth.fasta_bytes(synth.mutate(g0, rate, 20240100 + i))

        t0 = time.perf_counter()
        queries = [assembly(i) if i in mine else b"" for i in range(n)]     # a rank only materialises its own share
        gen_s = time.perf_counter() - t0
        warm = [queries[i] for i in sorted(mine)[:64]]
        classify_batch(eng, warm, refs, k=K, s=S)                            # warm-up (fills the pool of 64 sketch buffers a dist batch holds)
        sw_sk, sw_di = Stopwatch("sketch_bytes"), Stopwatch("dist")
        if dist is not None:                                                 # ... and NCCL's point-to-point channels (first gather: ~1 s)
            dist.gather_object({}, [None] * world if rank == 0 else None, dst=0)
        l0 = eng.launch_count()
        barrier()
        t0 = time.perf_counter()
        rows = classify_batch(eng, queries, refs, k=K, s=S, rank=rank, world=world)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        sw_sk.restore(); sw_di.restore()
        rec = None
        if rank == 0:
            from oracle import oracle
            ok, checked = True, []
            for qi in sorted({0, n - 1, *np.random.default_rng(1).integers(0, n, max(0, args.check - 2)).tolist()}):
                osk = oracle.sketch_parallel([assembly(qi)], K, S, 1, threads=len(os.sched_getaffinity(0)))
                for ri, (rh, rl) in enumerate(refs_sk):
                    d, p, cm, dn = oracle.dist(rh, osk.hashes, S, K, rl, osk.total_len)
                    row = rows[qi * 24 + ri]
                    good = (row.common, row.denom) == (cm, dn) and abs(row.distance - d) <= 1e-12 * max(abs(d), 1e-300) and \
                        (abs(row.pvalue - p) <= 1e-12 * abs(p) or abs(row.pvalue - p) < 1e-300)
                    ok &= bool(good)
                checked.append(qi)
            mb = 12.4e6 * 1.0125 / 1e6
            rec = {"path": "c4_batch_classification", "n_gpus": world, "assemblies": n, "rows": len(rows), "seconds": dt,
                   "assemblies_per_s": n / dt, "ms_per_assembly_per_gpu": 1e3 * dt / (n / world), "fasta_MB_per_assembly": mb,
                   "Gbases_per_s": 12.4e6 * n / dt / 1e9, "gpu_launches_rank0": eng.launch_count() - l0,
                   "parity": {"checked_queries": checked, "rows_checked": 24 * len(checked), "equal": ok,
                              "against": "oracle.sketch_parallel + oracle.dist: common/denom bit for bit, distance and p-value 1e-12"},
                   "generation_s_rank0": gen_s, "rank0_calls": {"sketch_bytes": sw_sk.stats(), "dist": sw_di.stats()},
                   "includes": "H2D of the FASTA text from pageable host memory, K1/K2/count table, 24 dist rows per query, rows "
                               "gathered on rank 0 (gather_object); query q -> rank q mod P, references replicated"}
        if rec:
            emit(rec)
        refs.free()
        del queries
    if "c5" not in args.skip:
        t0 = time.perf_counter()
        sk = c5_sketches(args.sketches, S)
        sset = eng.load_sketch_set(sk, [12_400_000] * len(sk), K, S)
        gen_s = time.perf_counter() - t0
        list(all_pairs(eng, sset, min(args.tile, len(sk)), tile=args.tile))     # warm-up: one full tile (sizes the result buffers)
        l0 = eng.launch_count()
        sw_dp = Stopwatch("dist_pairs", device=True)
        barrier()
        t0 = time.perf_counter()
        pairs, common_sum, sample = 0, 0, []
        for r in all_pairs(eng, sset, len(sk), tile=args.tile, rank=rank, world=world):
            pairs += r.common.size
            common_sum += int(r.common.sum())
            if len(sample) < args.check:
                sample.append(r)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        sw_dp.restore()
        tot = np.array([pairs, common_sum], np.float64)
        if dist is not None:
            import torch
            t = torch.tensor(tot, device="cuda")
            dist.all_reduce(t)
            tot = t.cpu().numpy()
        # this rank's sample of tiles against the oracle (16 pairs per tile: corners, diagonal, random)
        from oracle import oracle
        ok = True
        rng = np.random.default_rng(7 + rank)
        n_checked = 0
        for r in sample:
            t = r.tile
            pts = {(0, 0), (t.ni - 1, t.nj - 1), (0, t.nj - 1), (t.ni - 1, 0)} | {(int(rng.integers(0, t.ni)), int(rng.integers(0, t.nj))) for _ in range(12)}
            for a, b in pts:
                d, p, cm, dn = oracle.dist(sk[t.i0 + a], sk[t.j0 + b], S, K, 12_400_000, 12_400_000)
                ok &= bool((int(r.common[a, b]), int(r.denom[a, b])) == (cm, dn) and abs(r.distance[a, b] - d) <= 1e-12 * max(abs(d), 1e-300)
                           and (abs(r.pvalue[a, b] - p) <= 1e-12 * abs(p) or abs(r.pvalue[a, b] - p) < 1e-300))
                n_checked += 1
        if dist is not None:
            import torch
            t = torch.tensor([1 if ok else 0, n_checked], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = bool(t[0].item())
        if rank == 0:
            pairs_all, common_all = int(tot[0]), int(tot[1])
            steps = pairs_all * S + common_all      # merge steps Mash's loop would run: until the union reaches s, i.e. s + common
            full = 50_000 * 50_001 // 2
            emit({"path": "c5_all_vs_all_dist", "n_gpus": world, "sketches": len(sk), "hashes_per_sketch": S, "tile": args.tile,
                  "tiles": len(tile_plan(len(sk), args.tile, world)), "pairs": pairs_all, "seconds": dt, "pairs_per_s": pairs_all / dt,
                  "pairs_per_s_per_gpu": pairs_all / dt / world, "merge_steps_per_s": steps / dt,
                  "algorithmic_GB_per_s_per_gpu": steps * 8 / dt / 1e9 / world,
                  "extrapolated_seconds_for_50000_sketches": full / (pairs_all / dt),
                  "extrapolation": "1.25e9 unordered pairs of BASELINE configs[4] at the measured pairs/s (same s, same cluster structure; "
                                   "the 20 GB sketch set fits every GPU)",
                  "parity": {"pairs_checked_per_rank": n_checked, "equal": ok,
                             "against": "oracle.dist (sequential capped merge): common/denom bit for bit, distance and p-value 1e-12"},
                  "gpu_launches_rank0": eng.launch_count() - l0, "generation_and_upload_s_rank0": gen_s,
                  "rank0_calls": {"dist_pairs": sw_dp.stats()},
                  "device_pairs_per_s_rank0": pairs / (sw_dp.stats()["device_ms_total"] / 1e3),
                  "includes": "upper-triangle tiles dealt round-robin over the ranks, four [tile, tile] result arrays copied back to the host per tile"})
        sset.free()
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
