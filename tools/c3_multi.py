#!/usr/bin/env python
"""BASELINE config 3: hall-500k, Octree(7,32), one Shoot per ray, the ray batch block-sharded over the GPUs
of one box (torchrun, one rank per GPU), X_Event rows gathered to rank 0 over NCCL/NVLink.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/c3_multi.py --rays-total 100000000

Strong scaling: the total is fixed, every rank generates and shoots only its block.  Timing: CUDA events on
each rank around Shoot + gather, max over ranks; geometry build excluded (done once per model).
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hare_b200 as hb  # noqa: E402
from hare_b200 import dist as hd  # noqa: E402
from hare_b200._lib import check, lib  # noqa: E402
from hare_b200.harness import meshes, rays_from_sources  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays-total", type=int, default=100_000_000)
    ap.add_argument("--mesh", default="500k"); ap.add_argument("--part", default="octree", choices=["octree", "voxel", "kdtree"])
    ap.add_argument("--args", type=int, nargs="+", default=[7, 32]); ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    hb.init([local])
    mesh = meshes.hall(a.mesh)
    T = hb.Topology.from_mesh(mesh)
    part = {"voxel": hb.Voxel_Grid, "octree": hb.Octree, "kdtree": hb.KDTree}[a.part]([T], *a.args)
    lo, hi = hd.shard_range(a.rays_total, rank, world)
    N = hi - lo
    o, d = rays_from_sources(N, meshes.sources(8), stream=3, first=lo)
    o_d = torch.from_numpy(o).to(dev); d_d = torch.from_numpy(d).to(dev)
    del o, d
    t = torch.empty(N, dtype=torch.float64, device=dev); xyz = torch.empty((N, 3), dtype=torch.float64, device=dev)
    pid = torch.empty(N, dtype=torch.int32, device=dev); uv = torch.empty((N, 2), dtype=torch.float64, device=dev)
    sizes = [hd.shard_range(a.rays_total, r, world)[1] - hd.shard_range(a.rays_total, r, world)[0] for r in range(world)]
    L = lib()

    def step():
        check(L.hare_shoot_batch_device(part._h, o_d.data_ptr(), d_d.data_ptr(), None, None, None, N, t.data_ptr(), xyz.data_ptr(),
                                        pid.data_ptr(), uv.data_ptr(), None, None, C.c_void_p(1)), "shoot")
        if world > 1:
            g = [hd.gather_rows(x, 0, sizes) for x in (pid, t, xyz, uv)]
            return g
        return None
    step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        g = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = hd.max_over_ranks(e0.elapsed_time(e1), dev) / a.steps
    hits = hd.sum_over_ranks(int((pid >= 0).sum().item()), dev)
    if rank == 0:
        ok = True
        if world > 1:
            ok = int((g[0] >= 0).sum().item()) == hits and g[0].shape[0] == a.rays_total
        print(json.dumps({"workload": f"C3: hall-{a.mesh} ({mesh.P} polygons), {a.part}{tuple(a.args)}, {a.rays_total} rays sharded over {world} GPU(s), X_Event gather to rank 0 (NCCL)",
                          "n_gpus": world, "ms_per_step": ms, "Mrays_per_s": a.rays_total / ms / 1e3, "hit_fraction": hits / a.rays_total,
                          "gather_bytes_per_step": int(a.rays_total * 52 * (world - 1) / max(world, 1)), "gathered_rows_consistent": ok}))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
