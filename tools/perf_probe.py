#!/usr/bin/env python
"""Device-resident timing of single-Shoot batches on the other BASELINE configs (C3, C4, C5).
Not the bench contract (bench.py is): a developer probe whose output feeds profiles/ and DESIGN.md.

    python tools/perf_probe.py --mesh 500k --part octree --args 7 32 --rays 4000000
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hare_b200 as hb  # noqa: E402
from hare_b200._lib import check, lib  # noqa: E402
from hare_b200.harness import meshes, rays_from_sources  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", default="500k")
    ap.add_argument("--part", default="octree", choices=["voxel", "octree", "kdtree"])
    ap.add_argument("--args", type=int, nargs="+", default=[7, 32])
    ap.add_argument("--rays", type=int, default=4_000_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--sources", type=int, default=8)
    ap.add_argument("--check", type=int, default=0, help="compare the first K rays with the CPU oracle")
    a = ap.parse_args()
    hb.init([0])
    torch.cuda.set_device(0)
    t0 = time.perf_counter(); mesh = meshes.hall(a.mesh); t_mesh = time.perf_counter() - t0
    t0 = time.perf_counter(); T = hb.Topology.from_mesh(mesh); t_topo = time.perf_counter() - t0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    part = {"voxel": hb.Voxel_Grid, "octree": hb.Octree, "kdtree": hb.KDTree}[a.part]([T], *a.args)
    torch.cuda.synchronize(); t_build = time.perf_counter() - t0
    info = part.info()
    o, d = rays_from_sources(a.rays, meshes.sources(a.sources), stream=3)
    dev = torch.device("cuda", 0)
    o_d = torch.from_numpy(o).to(dev); d_d = torch.from_numpy(d).to(dev)
    N = a.rays
    t = torch.empty(N, dtype=torch.float64, device=dev); xyz = torch.empty((N, 3), dtype=torch.float64, device=dev)
    pid = torch.empty(N, dtype=torch.int32, device=dev); uv = torch.empty((N, 2), dtype=torch.float64, device=dev)
    cnt = torch.zeros(4, dtype=torch.int64, device=dev)
    L = lib()

    def run(counters=False):
        check(L.hare_shoot_batch_device(part._h, o_d.data_ptr(), d_d.data_ptr(), None, None, None, N, t.data_ptr(), xyz.data_ptr(),
                                        pid.data_ptr(), uv.data_ptr(), None, cnt.data_ptr() if counters else None, C.c_void_p(1)), "shoot")
    run(True); torch.cuda.synchronize()
    c = cnt.cpu().numpy() / N
    best = 1e9
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    hit = float((pid >= 0).float().mean())
    print(f"{a.mesh} P={mesh.P} {a.part}{tuple(a.args)} info={info} mesh {t_mesh:.1f}s topo {t_topo:.1f}s build {t_build:.2f}s | "
          f"{N} rays {best:.1f} ms = {N / best / 1e3:.1f} Mrays/s hit={hit:.4f} per-ray cells={c[0]:.1f} entries={c[1]:.1f} tests={c[2]:.1f}")
    if a.check:
        from oracle import hare_oracle as ho
        To = ho.Topology.from_mesh(mesh)
        if a.part == "voxel":
            op = ho.Voxel_Grid(To, a.args[0], "fast")
        elif a.part == "octree":
            op = ho.Octree(To, *a.args)
        else:
            op = ho.KDTree(To, *a.args)
        K = a.check
        t0 = time.perf_counter(); ref = op.Shoot(o[:K], d[:K], nthreads=os.cpu_count()); dt = time.perf_counter() - t0
        same_t = np.array_equal(ref["t"], t[:K].cpu().numpy()); diff = int((ref["poly_id"] != pid[:K].cpu().numpy()).sum())
        print(f"  oracle {K} rays in {dt:.2f}s ({K / dt / 1e6:.3f} Mrays/s, {os.cpu_count()} threads) counters/ray={ref['counters'][:3] / K}  t identical={same_t} poly_id mismatches={diff}")


if __name__ == "__main__":
    main()
