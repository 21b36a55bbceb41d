#!/usr/bin/env python
"""C5: Voxel_Grid build (AABB_Tri_Int binning) on the GPU vs the CPU restatement.

    python tools/build_probe.py --mesh 2m --domain 256 [--cpu hier|fast|none]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hare_b200 as hb  # noqa: E402
from hare_b200.harness import meshes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", default="2m"); ap.add_argument("--domain", type=int, default=256)
    ap.add_argument("--cpu", default="fast", choices=["hier", "fast", "none"]); ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    hb.init([0]); torch.cuda.set_device(0)
    mesh = meshes.hall(a.mesh)
    T = hb.Topology.from_mesh(mesh)
    best = 1e9
    for _ in range(a.reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        g = hb.Voxel_Grid([T], a.domain)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    obox, vd, ct, n = g.info()
    ncells = a.domain ** 3
    bytes_alg = 2 * mesh.P * 128 + 12 * ncells + 4 * n
    print(f"{a.mesh} P={mesh.P} Voxel_Grid {a.domain}^3: GPU build {best * 1e3:.1f} ms (host wall, incl. allocation) pairs={n} "
          f"= {mesh.P / best / 1e6:.1f} Mpolys/s, {n / best / 1e6:.1f} Mpairs/s, algorithmic {bytes_alg / 1e9:.2f} GB -> {bytes_alg / best / 1e9:.0f} GB/s")
    if a.cpu != "none":
        from oracle import hare_oracle as ho
        To = ho.Topology.from_mesh(mesh)
        t0 = time.perf_counter()
        if a.cpu == "hier":
            og = ho.Voxel_Grid(To, int(round(np.log2(a.domain))), "hier", avg_polys=0, nthreads=os.cpu_count())
        else:
            og = ho.Voxel_Grid(To, a.domain, "fast")
        dt = time.perf_counter() - t0
        off, pol = g.csr(); ooff, opol = og.csr()
        same = np.array_equal(off, ooff) and np.array_equal(pol, opol)
        print(f"  CPU {a.cpu} build {dt:.2f} s ({os.cpu_count()} threads for hier, 1 for fast) -> GPU is {dt / best:.0f}x; CSR identical: {same}"
              + ("" if same else f" (cells differing: {int((np.diff(off.astype(np.int64)) != np.diff(ooff.astype(np.int64))).sum())})"))


if __name__ == "__main__":
    main()
