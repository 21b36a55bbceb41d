"""world_size-2 gloo test of the multi-rank path's host logic (hare_b200/dist.py): block sharding of
the ray batch, per-rank Shoot, gather of X_Event rows onto rank 0 in rank order, max/sum reductions.
The per-rank compute here is the CPU oracle (there is no GPU in this test); on the GPU box the same
helpers run over NCCL inside bench.py."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, ragged, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hare_b200 import dist as hd
    from hare_b200.harness import meshes, rays_from_sources
    from oracle import hare_oracle as ho
    mesh = meshes.hall("tiny")
    o, d = rays_from_sources(n, meshes.sources(4), stream=13)      # every rank generates the same global batch
    part = ho.Voxel_Grid(ho.Topology.from_mesh(mesh), 6, "fast")   # geometry replicated on every rank
    lo, hi = hd.shard_range(n, rank, world)
    r = part.Shoot(o[lo:hi], d[lo:hi], ray_id=np.arange(lo + 1, hi + 1, dtype=np.int32))
    sizes = [hd.shard_range(n, k, world)[1] - hd.shard_range(n, k, world)[0] for k in range(world)]
    pid = hd.gather_rows(torch.from_numpy(r["poly_id"]), 0, sizes if ragged else None)
    t = hd.gather_rows(torch.from_numpy(r["t"]), 0, sizes if ragged else None)
    xyz = hd.gather_rows(torch.from_numpy(r["xyz"]), 0, sizes if ragged else None)
    tmax = hd.max_over_ranks(rank + 1.5, "cpu"); tot = hd.sum_over_ranks(hi - lo, "cpu")
    if rank == 0:
        ref = part.Shoot(o, d)
        ok = (np.array_equal(pid.numpy(), ref["poly_id"]) and np.array_equal(t.numpy(), ref["t"]) and np.array_equal(xyz.numpy(), ref["xyz"])
              and tmax == world + 0.5 and tot == n)
        q.put(bool(ok))
    else:
        assert pid is None and t is None
    dist.barrier()
    dist.destroy_process_group()


def _run(n, ragged):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, ragged, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_two_rank_shard_and_gather_equal_blocks():
    _run(4000, ragged=False)


def test_two_rank_shard_and_gather_ragged_blocks():
    _run(4001, ragged=True)


def test_shard_range_covers_everything():
    from hare_b200.dist import shard_range
    for n in (0, 1, 7, 100, 12345):
        for w in (1, 2, 3, 8):
            blocks = [shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1


def _worker_shared(rank, world, port, n, q):
    """bench.py's strong-scaling plumbing: block shards of ONE global batch (bench.shard), every rank writing its X_Event rows at its
    own offset into arrays owned by rank 0 (bench.SharedEvents: the host-side twin of dist.PeerResults, which needs GPUs)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import hare_b200 as hb
    from hare_b200.harness import meshes, rays_from_sources
    from oracle import hare_oracle as ho
    cfg = dict(bench.CONFIGS["C3"], name="C3", rays=n)
    lo, hi = bench.shard(cfg, rank, world)
    assert bench.total_rays(cfg, world) == n
    x = bench.Ctx(); x.rank, x.world, x.dist, x.L = rank, world, dist, hb.lib()
    mesh = meshes.hall("tiny")
    part = ho.Octree(ho.Topology.from_mesh(mesh), 3, 4)
    o, d = rays_from_sources(hi - lo, meshes.sources(8), stream=3, first=lo)       # every rank generates only its own block
    r = part.Shoot(o, d)
    ev = bench.SharedEvents(x, n, lo, hi - lo)
    t, xyz, pid, uv = ev.rows()
    t[:] = r["t"]; xyz[:] = r["xyz"]; pid[:] = r["poly_id"]; uv[:] = r["uv"]
    dist.barrier()
    if rank == 0:
        og, dg = rays_from_sources(n, meshes.sources(8), stream=3)
        ref = part.Shoot(og, dg)
        T, X, P, U = ev.whole()
        q.put(bool(np.array_equal(P, ref["poly_id"]) and np.array_equal(T, ref["t"]) and np.array_equal(X, ref["xyz"]) and np.array_equal(U, ref["uv"])))
    dist.barrier()
    ev.close()
    dist.destroy_process_group()


def test_two_rank_rows_land_in_one_shared_array():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_shared, args=(r, 2, port, 5001, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
