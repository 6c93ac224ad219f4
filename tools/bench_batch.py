#!/usr/bin/env python
"""Throughput of the two batch paths (BASELINE.json configs[3] and configs[4]) on one GPU -- secondary
measurements beside bench.py's headline metric.  Prints one JSON line per path.

  c4: batch classification -- N synthetic 12.4 Mb assemblies (G0 mutated at U(0, 2 %)), each sketched from HOST
      memory (H2D inside the timed region) and distanced against 24 references; assemblies/s.
  c5: all-vs-all dist -- T x T tiles of synthetic 50,000-hash sketches (clusters sharing a base set, SURVEY.md
      section 8d); pairs/s and merge steps/s (the oracle's loop count model: <= 2 s steps per pair).
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from auriclass_b200 import synth  # noqa: E402
from auriclass_b200._lib import AC_MODE_FASTA  # noqa: E402
from auriclass_b200.batch import all_pairs, classify_batch  # noqa: E402
from auriclass_b200.mash_gpu import MashGPU  # noqa: E402

K, S = 27, 50000


def c5_sketches(n, s, seed=20240500, clusters=8):
    rng = np.random.Generator(np.random.PCG64(seed))
    bases = [np.unique(rng.integers(0, 1 << 63, int(1.2 * s), dtype=np.uint64)) for _ in range(clusters)]
    out = []
    for i in range(n):
        h = bases[i % clusters].copy()
        rep = rng.random(h.size) < rng.uniform(0, 0.3)
        h[rep] = rng.integers(0, 1 << 63, int(rep.sum()), dtype=np.uint64)
        out.append(np.unique(h)[:s])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--assemblies", type=int, default=24)
    ap.add_argument("--sketches", type=int, default=768)
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--skip", default="")
    args = ap.parse_args()
    eng = MashGPU(0)
    if "c4" not in args.skip:
        g0 = synth.genome(20240001)
        refs_sk = []
        for i in range(24):
            fa = synth.fasta_bytes(synth.mutate(g0, 0.0005 + 0.0004 * i, 20240010 + i))
            sk = eng.sketch_bytes([fa], k=K, s=S, m=1, mode=AC_MODE_FASTA)
            refs_sk.append((sk.hashes()[0], sk.length))
            sk.free()
        refs = eng.load_sketch_set([h for h, _ in refs_sk], [l for _, l in refs_sk], K, S, names=[f"ref{i}" for i in range(24)])
        rng = np.random.Generator(np.random.PCG64(20240100))
        queries = [synth.fasta_bytes(synth.mutate(g0, float(rng.uniform(0, 0.02)), 20240100 + i)) for i in range(args.assemblies)]
        classify_batch(eng, queries, refs, k=K, s=S)              # warm-up (fills the sketch-buffer pool)
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        rows = classify_batch(eng, queries, refs, k=K, s=S)
        dt = time.perf_counter() - t0
        mb = sum(len(q) for q in queries) / 1e6
        print(json.dumps({"path": "c4_batch_classification", "assemblies": len(queries), "rows": len(rows), "seconds": dt,
                          "assemblies_per_s": len(queries) / dt, "ms_per_assembly": 1e3 * dt / len(queries),
                          "fasta_MB_per_assembly": mb / len(queries), "Gbases_per_s": 12.4e6 * len(queries) / dt / 1e9,
                          "gpu_launches": eng.launch_count() - l0,
                          "includes": "H2D of the FASTA text from host memory, K1/K2/count table, 24 dist rows back on the host"}), flush=True)
        refs.free()
    if "c5" not in args.skip:
        sk = c5_sketches(args.sketches, S)
        sset = eng.load_sketch_set(sk, [12_400_000] * len(sk), K, S)
        list(all_pairs(eng, sset, min(64, len(sk)), tile=64))     # warm-up
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        pairs, common_sum = 0, 0
        for r in all_pairs(eng, sset, len(sk), tile=args.tile):
            pairs += r.common.size
            common_sum += int(r.common.sum())
        dt = time.perf_counter() - t0
        # merge steps Mash's loop would run per pair: until the union reaches s, i.e. s + common (SURVEY.md section 8d unit)
        steps = pairs * S + common_sum
        print(json.dumps({"path": "c5_all_vs_all_dist", "sketches": len(sk), "hashes_per_sketch": S, "tile": args.tile, "pairs": pairs,
                          "seconds": dt, "pairs_per_s": pairs / dt, "merge_steps_per_s": steps / dt,
                          "algorithmic_GB_per_s": steps * 8 / dt / 1e9, "gpu_launches": eng.launch_count() - l0,
                          "includes": "upper-triangle tiles, four [tile, tile] result arrays copied back to the host per tile"}), flush=True)
        sset.free()
    eng.close()


if __name__ == "__main__":
    main()
