"""rows / halo per rank of the shard planner's layout (host-only, no GPU):  python tools/plan_stats.py L world [chunks]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "exactdiagonalization.jl_b200"))
import edcuda as ed
from edcuda._lib import lib, check


def main():
    n, world = int(sys.argv[1]), int(sys.argv[2])
    chunks = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    hs, h = ed.models.xxz_chain(n) if hasattr(ed.models, "xxz_chain") else ed.models.heisenberg_chain(n)
    op = h if hasattr(h, "handle") else ed.Operator(h)
    t0 = time.time()
    rows, halo, pulls = [], [], []
    for r in range(world):
        counts = (C.c_int64 * 12)()
        check(lib.ed_shard_plan_describe(op.handle(), n, n // 2, 0, world, r, chunks, 0, counts, None, None, None, None, None, None))
        rows.append(counts[0]); halo.append(counts[1]); pulls.append(counts[4])
    dim = counts[7]
    rows, halo = np.array(rows, float), np.array(halo, float)
    print(f"L={n} world={world}: {time.time() - t0:.1f} s; rows max {rows.max() / 1e6:.2f}M (+{100 * (rows.max() * world / dim - 1):.1f} %), "
          f"halo max {halo.max() / 1e6:.2f}M mean {halo.mean() / 1e6:.2f}M, pulls max {max(pulls)}, "
          f"model ms: kernel {rows.max() * 14.5e-9:.2f} comm {0.3 + halo.max() * 21.3e-9:.2f}")
    print("  rows", (rows / 1e6).round(1).tolist())
    print("  halo", (halo / 1e6).round(1).tolist())


main()


def chunk_profile(n, world, chunks=8, rank=0):
    """rows launched and halo rows fetched per launch chunk of one rank"""
    hs, h = ed.models.xxz_chain(n)
    op = h if hasattr(h, "handle") else ed.Operator(h)
    counts = (C.c_int64 * 12)()
    args = (op.handle(), n, n // 2, 0, world, rank, chunks, 0)
    check(lib.ed_shard_plan_describe(*args, counts, None, None, None, None, None, None))
    n_tiles, n_pulls = counts[3], counts[4]
    tiles = np.zeros((max(n_tiles, 1), 4), dtype=np.int64)
    pulls = np.zeros((max(n_pulls, 1), 5), dtype=np.int64)
    check(lib.ed_shard_plan_describe(*args, counts, None, tiles.ctypes.data, pulls.ctypes.data, None, None, None))
    return tiles[:n_tiles], pulls[:n_pulls]
