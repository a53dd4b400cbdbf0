#!/usr/bin/env python
"""Owner arrays of the multi-GPU bench instances, computed here (build container) into gpurun_cache/ so that a
multi-GPU `gpurun` call does not pay N x the partitioner's host time with the GPUs idle.  bench.shared_partition
uses a cached array only when its key (hash of the instance's scopes, world size, method) matches; without the
cache it computes the same array on rank 0.  Usage: tools/precompute_partitions.py [c2 target c4 c3] [worlds]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pydcop_b200 import generators as G  # noqa: E402
from pydcop_b200.multigpu import resolve_owner  # noqa: E402


def main():
    which = [a for a in sys.argv[1:] if not a.isdigit()] or ["c2", "target", "c3", "c4"]
    worlds = [int(a) for a in sys.argv[1:] if a.isdigit()] or [2, 4, 8]
    os.makedirs(os.path.join(ROOT, "gpurun_cache"), exist_ok=True)
    for w in which:
        for world in worlds:
            inst = {"c2": lambda: G.config_c2(seed=0, n_vars=100_000 * world), "target": G.config_target,
                    "c3": G.config_c3, "c4": G.config_c4}[w]()
            path = bench._partition_cache_path(inst, world, "auto")
            if os.path.exists(path):
                print("have", path)
                continue
            t = time.perf_counter()
            owner = resolve_owner(inst, world, "auto")
            np.save(path, np.asarray(owner, dtype=np.int8))
            print(f"{w} x{world}: {time.perf_counter() - t:.1f}s -> {path}")


if __name__ == "__main__":
    main()
