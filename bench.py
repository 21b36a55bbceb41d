#!/usr/bin/env python
"""bench.py -- Mrays/s of closest-hit Shoot on the BASELINE.json configurations.

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path, default --config C3
    python bench.py --impl reference --gpus N --steps K ...       # the CPU restatement of the reference, same config

Configurations (`--config`):
    C3   (default) hall-500k, Octree(Model, 7, 32), 100 M rays in total, one Shoot each.  The batch is block-sharded over
         the N ranks (STRONG scaling); every rank's X_Event rows (poly_id, t, X_Point, u, v = 52 B) are delivered into rank 0's
         buffers over NVLink inside the timed region -- the path's only cross-GPU step (peer copies overlapped with the next
         step's traversal; --gather selects the alternatives).
    C2   hall-50k, Voxel_Grid 64^3, 10 M rays x 50-order specular chains per GPU (weak scaling).
    C4vg / C4kd   hall-2m, Voxel_Grid 256^3 / KDTree(24, 16), 100 M rays.
    C5   Voxel_Grid build: 2 M polygons -> 256^3 cell lists (count / scan / scatter).
    C1   shoebox, Voxel_Grid 10^3, 100 k rays: parity configuration; also prints the boundary's cliffs (single-ray Shoot
         latency, pageable vs page-locked hare_shoot_batch).
At N = 1 the default run appends short C2 / C4 / C5 / C1 lines under "other_configs" (--no-extras skips them).

One "step" = one pass of the hot path over the whole ray batch.  `value` is device-resident throughput (rays already in
HBM, CUDA events, max over ranks); `e2e` goes through the host-buffer C-ABI call (hare_shoot_batch / hare_reflect_chain,
page-locked host arrays) with the H2D / D2H copies inside the timed region.  Before a line is printed the GPU results of
the first rays are compared bit for bit with the oracle's (the same run that is timed as cpu_baseline).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PEAKS_FALLBACK_HBM = 6650.0   # GB/s, B200_PROFILING.md fallback
METRIC = "Mrays/s closest-hit Shoot"

CONFIGS = {
    # kind, mesh level, partition, ctor args, rays, ray stream, sources, scaling, CPU sample cap, parity sample floor
    "C1": dict(kind="shoot", mesh="shoebox", part="Voxel_Grid", args=(10,), rays=100_000, stream=1, nsrc=1, scaling="strong", cpu_cap=100_000, cpu_min=100_000),
    "C2": dict(kind="chain", mesh="50k", part="Voxel_Grid", args=(64,), rays=10_000_000, order=50, stream=2, nsrc=4, scaling="weak", cpu_cap=400_000, cpu_min=20_000),
    "C3": dict(kind="shoot", mesh="500k", part="Octree", args=(7, 32), rays=100_000_000, stream=3, nsrc=8, scaling="strong", cpu_cap=4_000_000, cpu_min=200_000),
    "C4vg": dict(kind="shoot", mesh="2m", part="Voxel_Grid", args=(256,), rays=100_000_000, stream=4, nsrc=8, scaling="strong", cpu_cap=4_000_000, cpu_min=200_000),
    "C4kd": dict(kind="shoot", mesh="2m", part="KDTree", args=(24, 16), rays=100_000_000, stream=4, nsrc=8, scaling="strong", cpu_cap=20_000, cpu_min=2_000),
    "C5": dict(kind="build", mesh="2m", part="Voxel_Grid", args=(256,), scaling="strong"),
}
HEADER_BYTES = {"Voxel_Grid": 8, "Octree": 64, "KDTree": 16}     # h of SURVEY.md 8(d)
EVENT_BYTES = {"Voxel_Grid": 36, "Octree": 52, "KDTree": 52}    # E


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--rays", type=int, default=None, help="override the configuration's ray count (C2: per GPU; others: total)")
    ap.add_argument("--order", type=int, default=None)
    ap.add_argument("--mesh", default=None)
    ap.add_argument("--args", type=int, nargs="+", default=None, help="override the partition's constructor arguments")
    ap.add_argument("--part", default=None, choices=["Voxel_Grid", "Octree", "KDTree"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="profiling runs: skip the CPU leg (and with it the parity gate)")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the host-buffer leg")
    ap.add_argument("--no-extras", action="store_true", help="do not append the short lines of the other configurations at N = 1")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--ray-order", default="source-major", choices=["source-major", "interleaved"],
                    help="which source a ray of the workload starts at: 'source-major' = one source after the other, as a caller that loops over its "
                         "sources produces them (default); 'interleaved' = round-robin over the sources")
    ap.add_argument("--presort", action="store_true", help="experiment: hand the rays over sorted by (origin, direction cell) instead of in generator order")
    ap.add_argument("--gather", default="peer", choices=["peer", "peer-store", "nccl"],
                    help="N > 1: 'peer' = a step's rows are copied into rank 0's buffers over NVLink (copy engines) while the next step traverses "
                         "(default); 'peer-store' = the kernels store their rows into rank 0's buffers directly; 'nccl' = NCCL gather after the kernel")
    return ap.parse_args()


def cfg_of(args, name=None):
    c = dict(CONFIGS[name or args.config])
    c["name"] = name or args.config
    if name is None:   # overrides only apply to the configuration named on the command line
        for k in ("rays", "order", "mesh", "part"):
            if getattr(args, k) is not None:
                c[k] = getattr(args, k)
        if args.args is not None:
            c["args"] = tuple(args.args)
    c["ray_order"] = getattr(args, "ray_order", "source-major")
    return c


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return PEAKS_FALLBACK_HBM, "fallback (B200_PROFILING.md)"


def ncu_actual(cfg):
    """Counters of the dominant kernel from the committed `ncu --set full` capture of this configuration at HEAD
    (profiles/r2_ncu_<config>.json, written by tools/ncu_kernel_json.py), or None."""
    p = os.path.join(ROOT, "profiles", f"r2_ncu_{cfg['name']}.json")
    try:
        j = json.load(open(p))
    except Exception:
        return None
    c = j.get("config", {})
    if (c.get("mesh"), c.get("part"), tuple(c.get("args", ()))) != (cfg["mesh"], cfg["part"], tuple(cfg["args"])):
        return None
    return j


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------------------------
def get_mesh(cfg):
    from hare_b200.harness import meshes
    return meshes.shoebox() if cfg["mesh"] == "shoebox" else meshes.hall(cfg["mesh"])


def get_sources(cfg):
    from hare_b200.harness import meshes
    return np.array([[5.0, 3.5, 1.5]]) if cfg["mesh"] == "shoebox" else meshes.sources(cfg["nsrc"])


def gen_rays(cfg, world, n, first, threads=None, out=None):
    """Rays first .. first + n - 1 of the configuration's workload (global ray numbers)."""
    from hare_b200.harness import rays_from_sources
    return rays_from_sources(n, get_sources(cfg), stream=cfg["stream"], first=first, threads=threads, out=out,
                             order=cfg["ray_order"], total=total_rays(cfg, world))


def take(a, idx):
    """Rows idx of a numpy array or a device tensor, as numpy."""
    if isinstance(a, np.ndarray):
        return a[idx]
    import torch
    return a[torch.from_numpy(idx).to(a.device)].cpu().numpy()


def shard(cfg, rank, world):
    """[lo, hi) of the global ray sequence this rank shoots."""
    if cfg["scaling"] == "weak":
        return rank * cfg["rays"], (rank + 1) * cfg["rays"]
    return (cfg["rays"] * rank) // world, (cfg["rays"] * (rank + 1)) // world


def total_rays(cfg, world):
    return cfg["rays"] * world if cfg["scaling"] == "weak" else cfg["rays"]


def workload_config(cfg, mesh, world, extra=None):
    ctor = f"{cfg['part']}(Model, {', '.join(map(str, cfg['args']))})"
    if cfg["kind"] == "build":
        w = f"{cfg['name']}: Voxel_Grid build, procedural hall-{cfg['mesh']} ({mesh.P} polygons) -> {cfg['args'][0]}^3 cell lists"
    elif cfg["kind"] == "chain":
        w = (f"{cfg['name']}: procedural auditorium hall-{cfg['mesh']} ({mesh.P} polygons), {ctor}, "
             f"{cfg['rays']} rays x {cfg['order']}-order specular chains per GPU")
    else:
        w = (f"{cfg['name']}: procedural hall {mesh.name} ({mesh.P} polygons), {ctor}, {total_rays(cfg, world)} rays in total from {cfg['nsrc']} source(s) "
             f"({cfg['ray_order']} order), one closest-hit Shoot each, block-sharded over {world} GPU(s)" + (", X_Event rows delivered to rank 0 over NVLink inside the timed region" if world > 1 else ""))
    c = {"workload": w, "partition": cfg["part"], "ctor_args": list(cfg["args"]), "polygons": mesh.P,
         "rays_total": total_rays(cfg, world) if cfg["kind"] != "build" else None,
         "ray_order": cfg["ray_order"] if cfg["kind"] != "build" else None,
         "l2": "ray inputs and event outputs stream through HBM every step (far larger than the 126 MB L2); geometry is re-read from L2/HBM as the walk needs it"}
    if cfg["kind"] == "chain":
        c["order"] = cfg["order"]
    if extra:
        c.update(extra)
    return c


def oracle_partition(cfg, mesh, nthreads):
    from oracle import hare_oracle as ho
    To = ho.Topology.from_mesh(mesh)
    if cfg["part"] == "Voxel_Grid":
        return ho.Voxel_Grid(To, cfg["args"][0], mode="fast", nthreads=nthreads)
    return getattr(ho, cfg["part"])(To, *cfg["args"])


def algorithmic_bytes(counters, shots, part):
    """B = 56 + E + h*C + 4*L + 128*T per Shoot."""
    C_, L_, T_ = (float(counters[k]) / max(1, shots) for k in range(3))
    return 56 + EVENT_BYTES[part] + HEADER_BYTES[part] * C_ + 4 * L_ + 128 * T_, dict(cells_or_nodes=C_, entries=L_, tests=T_)


def cpu_leg(cfg, mesh, o, d, seconds, nthreads):
    """Time the oracle (C++ restatement of the reference's CPU path) on a bounded sample of the batch: 64 evenly spaced runs of
    consecutive rays (every source is in it whatever the ray order), sized for about `seconds` of CPU time."""
    from hare_b200.harness import sample_blocks
    t0 = time.perf_counter()
    part = oracle_partition(cfg, mesh, nthreads)
    build_s = time.perf_counter() - t0
    n0 = int(min(len(o), max(256, cfg["cpu_min"] // 8)))
    if cfg["kind"] == "chain":
        run = lambda oo, dd: part.reflect_chain(oo, dd, cfg["order"], events=False, nthreads=nthreads)
        shots_of = lambda r: int(r["nshots"].sum())
    else:
        run = lambda oo, dd: part.Shoot(oo, dd, nthreads=nthreads)
        shots_of = lambda r: len(r["poly_id"])
    idx = sample_blocks(len(o), n0)
    oo, dd = np.ascontiguousarray(o[idx]), np.ascontiguousarray(d[idx])
    t0 = time.perf_counter(); r = run(oo, dd); dt = time.perf_counter() - t0
    per_ray = dt / max(1, len(idx))
    n = int(min(len(o), cfg["cpu_cap"], max(cfg["cpu_min"], seconds / max(per_ray, 1e-12))))
    idx = sample_blocks(len(o), n)
    oo, dd = np.ascontiguousarray(o[idx]), np.ascontiguousarray(d[idx])
    t0 = time.perf_counter(); r = run(oo, dd); dt = time.perf_counter() - t0
    shots = shots_of(r)
    return dict(mrays=shots / dt / 1e6, shots=shots, n=len(idx), idx=idx, seconds=dt, build_seconds=build_s, counters=r["counters"], result=r, part=part)


# --------------------------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement on all host threads
# --------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = cfg_of(args)
    cores = os.cpu_count() or 1
    mesh = get_mesh(cfg)
    world = args.gpus
    if cfg["kind"] == "build":
        return run_reference_build(args, cfg, mesh, cores)
    from hare_b200.harness import sample_rays
    per_step = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    n_gen = int(min(cfg["rays"], cfg["cpu_cap"]))
    # a bounded sample of the workload: 64 evenly spaced runs of consecutive rays (every source is in it)
    _, o, d = sample_rays(total_rays(cfg, world), n_gen, get_sources(cfg), stream=cfg["stream"], order=cfg["ray_order"])
    n_gen = len(o)
    part = oracle_partition(cfg, mesh, cores)
    chain = cfg["kind"] == "chain"
    picked = {}

    def pick(n):   # n of the sample's rays, again spread over all of it (cached: the copy stays out of the timed steps)
        if n not in picked:
            k = max(1, n_gen // max(1, n))
            picked[n] = (np.ascontiguousarray(o[::k][:n]), np.ascontiguousarray(d[::k][:n]))
        return picked[n]
    run = (lambda n: int(part.reflect_chain(*pick(n), cfg["order"], events=False, nthreads=cores)["nshots"].sum())) if chain else \
          (lambda n: len(part.Shoot(*pick(n), nthreads=cores)["poly_id"]))
    n0 = int(min(n_gen, max(256, cfg["cpu_min"] // 8)))
    t0 = time.perf_counter(); run(n0); dt = time.perf_counter() - t0
    n = int(min(n_gen, max(n0, per_step / (dt / n0))))
    for _ in range(args.warmup):
        run(n)
    t0 = time.perf_counter(); shots = 0
    for _ in range(args.steps):
        shots += run(n)
    dt = time.perf_counter() - t0
    val = shots / dt / 1e6
    sample = (f"{n} rays spread evenly over the workload" + (f" x {cfg['order']}-order chains" if chain else "") +
              f" per step ({shots // max(1, args.steps)} Shoots/step) of the {total_rays(cfg, world)}-ray workload")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, mesh, world),
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": f"C++ restatement of Hare's {cfg['part']}.Shoot CPU path (oracle/), std::thread over rays in contiguous chunks; no .NET runtime exists on this box",
    }))


def run_reference_build(args, cfg, mesh, cores):
    from oracle import hare_oracle as ho
    To = ho.Topology.from_mesh(mesh)
    dom = cfg["args"][0]
    md = int(round(np.log2(dom)))
    t0 = time.perf_counter()
    g = ho.Voxel_Grid(To, md, mode="hier", avg_polys=0, nthreads=cores)
    dt = time.perf_counter() - t0
    val = mesh.P / dt / 1e6
    print(json.dumps({
        "impl": "reference", "metric": "Mpolys/s Voxel_Grid build", "value": val, "unit": "Mpolys/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, mesh, 1),
        "cpu_baseline": {"value": val, "unit": "Mpolys/s", "cores": cores, "kind": "port",
                         "sample": f"one full hierarchical build Voxel_Grid(Model, MaxDomain={md}, Avg_polys=0) of all {mesh.P} polygons ({g.info()[3]} pairs)"},
        "e2e": {"value": val, "unit": "Mpolys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def make_ctx(args):
    import torch
    import torch.distributed as dist
    import hare_b200 as hb
    from hare_b200._lib import lib
    x = Ctx()
    x.world = int(os.environ.get("WORLD_SIZE", "1")); x.rank = int(os.environ.get("RANK", "0")); x.local = int(os.environ.get("LOCAL_RANK", "0"))
    if x.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(x.local)
    x.dev = torch.device("cuda", x.local)
    if x.world > 1:
        dist.init_process_group("nccl", device_id=x.dev)
    hb.init([x.local])
    x.L, x.hb, x.torch, x.dist = lib(), hb, torch, dist
    x.cores = max(1, (os.cpu_count() or 1) // (x.world if x.world > 1 else 1))
    return x


def cur_stream(torch):
    # torch's default stream is the legacy default stream (handle 0); the C ABI treats NULL as "the partition's own
    # stream", so name the legacy stream explicitly (cudaStreamLegacy == 0x1).
    s = torch.cuda.current_stream().cuda_stream
    return s if s else 1


def sync_all(x):
    if x.world > 1:
        x.dist.barrier()
    x.torch.cuda.synchronize()


def pinned(x, shape, dtype):
    """numpy array over page-locked host memory from the library (hare_host_alloc)."""
    n = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    from hare_b200._lib import check
    check(x.L.hare_host_alloc(max(n, 8), C.byref(p)), "hare_host_alloc")
    buf = (C.c_char * max(n, 8)).from_address(p.value)
    a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)
    x.pins.append(p)
    return a


def free_pins(x):
    for p in x.pins:
        x.L.hare_host_free(p)
    x.pins = []


class SharedEvents:
    """t, xyz, poly_id, uv for `n_total` rays in POSIX shared memory created by rank 0; rank r works on rows [lo, lo + n)."""
    FIELDS = (("t", (), np.float64), ("xyz", (3,), np.float64), ("poly_id", (), np.int32), ("uv", (2,), np.float64))

    def __init__(self, x, n_total, lo, n):
        from multiprocessing import shared_memory
        self.x, self.n_total, self.lo, self.n = x, n_total, lo, n
        names = [None] * len(self.FIELDS)
        self.seg = []
        if x.rank == 0:
            for k, (name, tail, dt) in enumerate(self.FIELDS):
                nbytes = n_total * int(np.prod(tail, dtype=np.int64)) * np.dtype(dt).itemsize
                sm = shared_memory.SharedMemory(create=True, size=max(nbytes, 8))
                self.seg.append(sm); names[k] = sm.name
        x.dist.broadcast_object_list(names, src=0)
        if x.rank != 0:
            self.seg = [shared_memory.SharedMemory(name=nm) for nm in names]
        self.arr = [np.ndarray((n_total,) + tail, dtype=dt, buffer=sm.buf) for sm, (name, tail, dt) in zip(self.seg, self.FIELDS)]
        self.registered = True
        self._reg = []
        for a in self.arr:
            row = a[lo:lo + n]
            if row.nbytes == 0:
                continue
            rc = x.L.hare_host_register(C.c_void_p(row.ctypes.data), row.nbytes)
            if rc != 0:
                self.registered = False
            else:
                self._reg.append(row.ctypes.data)

    def rows(self):
        return tuple(a[self.lo:self.lo + self.n] for a in self.arr)

    def whole(self):
        return self.arr

    def close(self):
        for p in self._reg:
            self.x.L.hare_host_unregister(C.c_void_p(p))
        self.arr = None
        for sm in self.seg:
            sm.close()
        self.x.dist.barrier()
        if self.x.rank == 0:
            for sm in self.seg:
                sm.unlink()


def reduce_max_sum(x, ms, count):
    torch, dist = x.torch, x.dist
    tms = torch.tensor([ms], dtype=torch.float64, device=x.dev); sh = torch.tensor([count], dtype=torch.int64, device=x.dev)
    if x.world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX); dist.all_reduce(sh, op=dist.ReduceOp.SUM)
    return float(tms.item()), int(sh.item())


def run_shoot_config(x, args, cfg, steps, warmup, cpu_seconds, do_e2e=True, do_cpu=True, gather="peer"):
    """Single-Shoot batch (C1 / C3 / C4) or reflection chains (C2).  Returns the JSON line (rank 0) or None."""
    torch, dist, L, hb = x.torch, x.dist, x.L, x.hb
    from hare_b200._lib import check
    from hare_b200.harness import rays_from_sources
    from hare_b200 import dist as hd
    chain = cfg["kind"] == "chain"
    order = cfg.get("order", 1)
    mesh = get_mesh(cfg)
    lo, hi = shard(cfg, x.rank, x.world)
    N = hi - lo
    x.pins = []
    # rays of this rank's block, generated straight into page-locked host arrays (the e2e leg shoots from them)
    o = pinned(x, (N, 3), np.float64); d = pinned(x, (N, 3), np.float64)
    gen_rays(cfg, x.world, N, lo, threads=x.cores, out=(o, d))
    if getattr(args, "presort", False):
        a = np.abs(d); face = np.argmax(a, axis=1); sgn = (np.take_along_axis(d, face[:, None], 1)[:, 0] < 0)
        u = np.take_along_axis(d, ((face + 1) % 3)[:, None], 1)[:, 0] / np.take_along_axis(a, face[:, None], 1)[:, 0]
        v = np.take_along_axis(d, ((face + 2) % 3)[:, None], 1)[:, 0] / np.take_along_axis(a, face[:, None], 1)[:, 0]
        ui = np.clip(((u + 1) * 32).astype(np.int64), 0, 63); vi = np.clip(((v + 1) * 32).astype(np.int64), 0, 63)
        mort = np.zeros(N, np.int64)
        for b in range(6):
            mort |= ((ui >> b) & 1) << (2 * b) | ((vi >> b) & 1) << (2 * b + 1)
        from hare_b200.harness import source_index
        src = source_index(np.arange(N) + lo, max(1, cfg["nsrc"]), cfg["ray_order"], total_rays(cfg, x.world))
        key = ((src * 6 + face * 2 + sgn) << 12) | mort
        perm = np.argsort(key, kind="stable")
        o[:] = o[perm]; d[:] = d[perm]
    cpu = None
    if x.rank == 0 and do_cpu:
        # the CPU leg runs first, on an otherwise idle host
        cpu = cpu_leg(cfg, mesh, o, d, cpu_seconds, os.cpu_count() or 1)
    if x.world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    T = hb.Topology.from_mesh(mesh)
    part = getattr(hb, cfg["part"])([T], *cfg["args"])
    build_s = time.perf_counter() - t0

    # ---------------- device-resident leg: rays already in HBM ---------------------------------------------------
    o_d = torch.from_numpy(o).to(x.dev); d_d = torch.from_numpy(d).to(x.dev)
    total = torch.zeros(1, dtype=torch.int64, device=x.dev)
    peer = None
    if chain:
        fin_o = torch.empty_like(o_d); fin_d = torch.empty_like(d_d); nshots = torch.empty(N, dtype=torch.int32, device=x.dev)
        # the per-bounce event streams (Poly_id, t: 12 B per Shoot) are written in the timed region: they are what a caller consumes
        ev_pid_d = torch.empty((N, order), dtype=torch.int32, device=x.dev); ev_t_d = torch.empty((N, order), dtype=torch.float64, device=x.dev)
        outs = [fin_o, fin_d, nshots]
    else:
        ntot = total_rays(cfg, x.world)
        if x.world > 1 and gather in ("peer", "peer-store"):
            try:
                peer = hd.PeerResults(ntot, x.local, dst=0)
            except Exception as e:   # no peer access between these devices: NCCL gather after the kernel instead
                if x.rank == 0:
                    print(f"# peer result buffers unavailable ({e}); using NCCL gather", file=sys.stderr)
                peer = None
            ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=x.dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                peer = None
        direct = peer is not None and (gather == "peer-store" or x.rank == 0)      # rank 0's own rows never need a copy
        nbuf = 1 if (peer is None or direct) else 2                                 # 'peer': double-buffered local rows
        if direct:
            pts = [peer.out_ptrs(lo)]
        else:
            bufs = [dict(t=torch.empty(N, dtype=torch.float64, device=x.dev), xyz=torch.empty((N, 3), dtype=torch.float64, device=x.dev),
                         poly_id=torch.empty(N, dtype=torch.int32, device=x.dev), uv=torch.empty((N, 2), dtype=torch.float64, device=x.dev)) for _ in range(nbuf)]
            pts = [(b["t"].data_ptr(), b["xyz"].data_ptr(), b["poly_id"].data_ptr(), b["uv"].data_ptr()) for b in bufs]
            t_d, xyz_d, pid_d, uv_d = bufs[0]["t"], bufs[0]["xyz"], bufs[0]["poly_id"], bufs[0]["uv"]
            outs = [pid_d, t_d, xyz_d, uv_d]
    sizes = [shard(cfg, r, x.world)[1] - shard(cfg, r, x.world)[0] for r in range(x.world)]
    gathered = None
    # 'peer': a rank's X_Event rows travel to rank 0 as four large peer copies (copy engines, NVLink; rank 0's buffers are mapped here
    # through CUDA IPC) on a second stream, WHILE the next step's batch is traversed -- local rows are double-buffered.  A step's rows
    # are complete on rank 0 one kernel later at the latest; the timed region ends only when the last step's copies have landed.
    # (Storing the rows from inside the kernel throttles it at 8 GPUs -- 52 B per ray as four small NVLink writes from seven peers
    # into one GPU: 51 vs 19 ms --, and cutting the batch into chunks to overlap within a step costs the persistent kernel 44 %.)
    copy_stream = torch.cuda.Stream(device=x.dev) if (peer is not None and not direct) else None
    if copy_stream is not None:
        remote = {k: peer.arrays[k].torch() for k in ("poly_id", "t", "xyz", "uv")}
        copied = [None] * nbuf
    step_no = [0]

    def kernel():
        stream = cur_stream(torch)
        if chain:
            check(L.hare_reflect_chain_device(part._h, o_d.data_ptr(), d_d.data_ptr(), N, order, ev_pid_d.data_ptr(), ev_t_d.data_ptr(), fin_o.data_ptr(),
                                              fin_d.data_ptr(), nshots.data_ptr(), total.data_ptr(), None, C.c_void_p(stream)), "hare_reflect_chain_device")
            return
        k = step_no[0] % nbuf
        if copy_stream is not None and copied[k] is not None:
            torch.cuda.current_stream().wait_event(copied[k])          # the copy that last read this buffer has finished
        pt, pxyz, ppid, puv = pts[k]
        check(L.hare_shoot_batch_device(part._h, o_d.data_ptr(), d_d.data_ptr(), None, None, None, N, C.c_void_p(pt), C.c_void_p(pxyz),
                                        C.c_void_p(ppid), C.c_void_p(puv), None, None, C.c_void_p(stream)), "hare_shoot_batch_device")

    def deliver():
        nonlocal gathered
        if x.world == 1:
            return
        if peer is not None:
            if copy_stream is not None:
                k = step_no[0] % nbuf
                done = torch.cuda.Event(); done.record()
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done)
                    for name in ("poly_id", "t", "xyz", "uv"):
                        remote[name][lo:lo + N].copy_(bufs[k][name], non_blocking=True)
                    copied[k] = torch.cuda.Event(); copied[k].record()
            elif gather == "peer-store":
                peer.fence()
            step_no[0] += 1
        else:
            gathered = [hd.gather_rows(a, 0, sizes) for a in outs]

    def drain():
        """End of a run of steps: every rank's last copies have landed on rank 0 (stream-ordered; the barrier follows in sync_all)."""
        if copy_stream is not None:
            torch.cuda.current_stream().wait_stream(copy_stream)
        if peer is not None:
            peer.fence()

    for _ in range(warmup):
        kernel(); deliver()
    drain()
    sync_all(x)
    total.zero_()
    torch.cuda.synchronize()
    sampler = ClockSampler(x.local) if x.rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = hb.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    sync_all(x)
    ev[0].record()
    for k in range(steps):
        kev[k][0].record(); kernel(); kev[k][1].record()   # inner events bracket the traversal kernel alone (roofline); outer ones the step
        deliver()
        if k == steps - 1:
            drain()
        ev[k + 1].record()
    sync_all(x)
    launches = hb.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms_total = ev[0].elapsed_time(ev[-1])
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    shots_rank = int(total.item()) if chain else N * steps
    ms_total, shots_all = reduce_max_sum(x, ms_total, shots_rank)
    kernel_ms_max, _ = reduce_max_sum(x, kernel_ms, 0)
    value = shots_all / (ms_total * 1e-3) / 1e6

    # results of the device leg on the host (rank 0: everything that was delivered to it)
    if chain:
        dev_res = dict(nshots=nshots.cpu().numpy(), o=fin_o)
    elif peer is not None:
        dev_res = {k: peer.arrays[k].torch() for k in ("poly_id", "t", "xyz", "uv")} if x.rank == 0 else None
    elif x.world > 1:
        dev_res = dict(zip(("poly_id", "t", "xyz", "uv"), gathered)) if x.rank == 0 else None
    else:
        dev_res = dict(poly_id=pid_d, t=t_d, xyz=xyz_d, uv=uv_d)

    # ---------------- end-to-end leg: host buffers through the public C-ABI call ------------------------------------
    e2e = None
    if do_e2e:
        if chain:
            fo_h = pinned(x, (N, 3), np.float64); fd_h = pinned(x, (N, 3), np.float64); ns_h = pinned(x, (N,), np.int32)
            tot = C.c_uint64()

            def step_host():
                check(L.hare_reflect_chain(part._h, o.ctypes.data, d.ctypes.data, N, order, None, None, fo_h.ctypes.data, fd_h.ctypes.data,
                                           ns_h.ctypes.data, C.byref(tot), None), "hare_reflect_chain")
                return tot.value
            h2d, d2h, api = 48, 52, "hare_reflect_chain (host buffers, page-locked)"
        else:
            shm = None
            if x.world > 1:
                # one host copy of the whole batch's events, shared by the ranks (POSIX shared memory created by rank 0): every rank
                # page-locks its own rows (hare_host_register) and hare_shoot_batch writes them there -- the gather ends in ONE host array
                shm = SharedEvents(x, total_rays(cfg, x.world), lo, N)
                t_h, xyz_h, pid_h, uv_h = shm.rows()
            else:
                t_h = pinned(x, (N,), np.float64); xyz_h = pinned(x, (N, 3), np.float64); pid_h = pinned(x, (N,), np.int32); uv_h = pinned(x, (N, 2), np.float64)

            def step_host():
                check(L.hare_shoot_batch(part._h, o.ctypes.data, d.ctypes.data, None, None, None, N, t_h.ctypes.data, xyz_h.ctypes.data,
                                         pid_h.ctypes.data, uv_h.ctypes.data, None, None), "hare_shoot_batch")
                return N
            h2d, d2h, api = 48, 52, "hare_shoot_batch (host buffers, page-locked)"
        for _ in range(min(warmup, 2)):
            step_host()
        sync_all(x)
        e_steps = max(1, min(steps, 3))
        t0 = time.perf_counter(); e_shots = 0
        for _ in range(e_steps):
            e_shots += step_host()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt, e_all = reduce_max_sum(x, dt, e_shots)
        e2e = {"value": e_all / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(total_rays(cfg, x.world) * h2d),
               "d2h_bytes_per_step": int(total_rays(cfg, x.world) * d2h), "steps": e_steps, "api": api,
               "note": (("every rank shoots its block from its own page-locked ray arrays into its rows of ONE shared host array "
                         "(POSIX shared memory owned by rank 0, rows page-locked with hare_host_register: " + ("yes" if shm.registered else "no, pageable") + ")")
                        if (x.world > 1 and not chain) else ("every rank shoots its own chains from and into its own page-locked host arrays" if x.world > 1 else None))}
        if chain:
            assert np.array_equal(ns_h, dev_res["nshots"]), "host-buffer and device-resident legs disagree"
            # the same call with the per-bounce event streams brought back to the host as well (12 B per Shoot more over PCIe)
            try:
                evp_h = pinned(x, (N, order), np.int32); evt_h = pinned(x, (N, order), np.float64)
                def step_host_ev():
                    check(L.hare_reflect_chain(part._h, o.ctypes.data, d.ctypes.data, N, order, evp_h.ctypes.data, evt_h.ctypes.data, fo_h.ctypes.data,
                                               fd_h.ctypes.data, ns_h.ctypes.data, C.byref(tot), None), "hare_reflect_chain")
                    return tot.value
                step_host_ev()
                sync_all(x)
                t0 = time.perf_counter(); ev_shots = step_host_ev(); torch.cuda.synchronize(); dt_ev = time.perf_counter() - t0
                dt_ev, ev_all = reduce_max_sum(x, dt_ev, ev_shots)
                e2e["with_event_streams"] = {"value": ev_all / dt_ev / 1e6, "unit": "Mrays/s", "d2h_bytes_per_step": int(total_rays(cfg, x.world) * (52 + 12 * order)),
                                             "note": "ev_poly_id + ev_t (N x order) copied back too"}
                assert np.array_equal(evp_h[:1000], ev_pid_d[:1000].cpu().numpy()), "event streams of the host-buffer and device-resident legs disagree"
            except MemoryError as e:
                e2e["with_event_streams"] = {"error": repr(e)}
        elif x.world == 1:
            assert np.array_equal(pid_h, dev_res["poly_id"].cpu().numpy()) and np.array_equal(t_h, dev_res["t"].cpu().numpy()), \
                "host-buffer and device-resident legs disagree"
        else:
            dist.barrier()
            if x.rank == 0:      # the shared host array holds every rank's rows: compare all of it with what the device leg delivered to rank 0
                full = shm.whole()
                assert np.array_equal(full[2], dev_res["poly_id"].cpu().numpy()) and np.array_equal(full[0], dev_res["t"].cpu().numpy()), \
                    "host-buffer (shared host array) and device-resident legs disagree"
            dist.barrier()
            shm.close()

    line = None
    if x.rank == 0:
        peak, peak_kind = peak_hbm()
        cores = os.cpu_count() or 1
        cpu_line, parity = None, None
        # GPU's own walk counters on a sample (nodes/cells entered, list entries scanned, exact tests)
        from hare_b200.harness import sample_blocks
        sidx = sample_blocks(N, 200_000)
        ns = len(sidx)
        so, sd = np.ascontiguousarray(o[sidx]), np.ascontiguousarray(d[sidx])
        if chain:
            r = part.Reflect_Chain(so, sd, order, events=False, counters=True)
            gpu_cnt, gpu_shots = r["counters"], r["total_shots"]
        else:
            r = part.Shoot_Batch(so, sd, counters=True)
            gpu_cnt, gpu_shots = r["counters"], ns
        gpu_bytes, gpu_avg = algorithmic_bytes(gpu_cnt, gpu_shots, cfg["part"])
        if cpu is not None:
            n = cpu["n"]
            ref = cpu["result"]
            cidx = cpu["idx"]
            if chain:
                ok = np.array_equal(ref["nshots"], take(dev_res["nshots"], cidx)) and np.array_equal(ref["o"], take(dev_res["o"], cidx))
            else:
                got = {k: take(dev_res[k], cidx) for k in ("poly_id", "t", "xyz", "uv")}
                ok = all(np.array_equal(got[k], ref[k]) for k in (("poly_id", "t", "xyz") if cfg["part"] == "Voxel_Grid" else ("poly_id", "t", "xyz", "uv")))
            assert ok, f"{cfg['name']}: GPU results differ from the oracle on the {n} sampled rays"
            parity = {"rays_compared": n, "bit_exact": True, "fields": "poly_id, t, X_Point" + ("" if cfg["part"] == "Voxel_Grid" else ", u, v")}
            if x.world > 1 and not chain:
                # rows that arrived from the OTHER ranks: the first rays of every remote block, re-generated here, against the oracle and
                # against this rank's own kernel
                m = 200 if cfg["part"] == "KDTree" else max(2_000, min(50_000, n // 8))    # the KDTree oracle is O(P) per ray
                checked = 0
                for r in range(1, x.world):
                    rlo = shard(cfg, r, x.world)[0]
                    ro, rd = gen_rays(cfg, x.world, m, rlo, threads=cores)
                    want = cpu["part"].Shoot(ro, rd, nthreads=cores)
                    for k in (("poly_id", "t", "xyz") if cfg["part"] == "Voxel_Grid" else ("poly_id", "t", "xyz", "uv")):
                        assert np.array_equal(dev_res[k][rlo:rlo + m].cpu().numpy(), want[k]), f"{cfg['name']}: rows delivered by rank {r} differ from the oracle ({k})"
                    checked += m
                parity["remote_rows_compared"] = checked
            ref_bytes, ref_avg = algorithmic_bytes(cpu["counters"], cpu["shots"], cfg["part"])
            cpu_line = {"value": cpu["mrays"], "unit": "Mrays/s", "cores": cores, "kind": "port",
                        "sample": f"{n} rays in 64 evenly spaced runs" + (f" x {order}-order chains" if chain else "") +
                                  f" ({cpu['shots']} Shoots, {cpu['seconds']:.1f} s) of the {total_rays(cfg, x.world)}-ray workload; "
                                  f"oracle partition build {cpu['build_seconds']:.1f} s not included"}
        else:
            ref_bytes, ref_avg = None, None
        # contract figure: reference-algorithm counts for Voxel_Grid / Octree, the pruned GPU count for KDTree
        if cfg["part"] == "KDTree" or ref_bytes is None:
            contract_bytes, contract_avg, contract_src = gpu_bytes, gpu_avg, "GPU walk counters (pruned)"
        else:
            contract_bytes, contract_avg, contract_src = ref_bytes, ref_avg, "oracle following the reference algorithm"
        shots_per_launch = shots_rank / steps
        achieved = shots_per_launch * contract_bytes / (kernel_ms * 1e-3) / 1e9
        achieved_gpu = shots_per_launch * gpu_bytes / (kernel_ms * 1e-3) / 1e9
        act = ncu_actual(cfg)
        kname = {"Voxel_Grid": "vg_wave_kernel", "Octree": "oct_wave_kernel", "KDTree": "kd_wave_kernel"}[cfg["part"]]
        roof = {"bound": (act or {}).get("bound", "issue/latency (see actual)"), "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_reference_bytes": achieved / peak,
                "frac_flag": "exceeds 1: the culls legitimately skip most of the reference's fetches; read achieved_gpu_counted / actual instead" if achieved / peak > 1.0 else None,
                "traffic": ((act or {}).get("traffic_bytes_per_ray") or 0) * shots_per_launch or None,
                "traffic_note": "ncu dram__bytes_read + dram__bytes_write per Shoot of the committed capture x the Shoots of one launch here" if act else None,
                "peak_kind": peak_kind, "kernel": kname, "kernel_ms": kernel_ms, "kernel_ms_max_over_ranks": kernel_ms_max,
                "bytes_per_shoot": contract_bytes, "per_shoot": contract_avg, "counted_by": contract_src,
                "formula": f"56 + {EVENT_BYTES[cfg['part']]} + {HEADER_BYTES[cfg['part']]}*nodes_or_cells + 4*entries + 128*tests (SURVEY.md 8(d))",
                "achieved_gpu_counted": achieved_gpu, "frac_gpu_counted": achieved_gpu / peak, "gpu_bytes_per_shoot": gpu_bytes, "gpu_per_shoot": gpu_avg,
                "actual": (act or {}).get("actual"), "actual_source": (f"profiles/r2_ncu_{cfg['name']}.json" if act else None)}
        line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": x.world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": workload_config(cfg, mesh, x.world, {"build_seconds_gpu": build_s}),
                "roofline": roof, "cpu_baseline": cpu_line, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "shots_per_step": shots_all // steps, "parity": parity,
                "result_delivery": None if x.world == 1 else (
                    ("each rank's rows go to rank 0's buffers (CUDA IPC mapping) as large peer copies over NVLink on a second stream while the NEXT step's batch "
                     "is traversed (double-buffered); the timed region ends when the last step's copies have landed"
                     if gather == "peer" else "peer stores from the traversal kernel straight into rank 0's buffers (CUDA IPC over NVLink) + 4-byte NCCL fence")
                    if peer is not None else "NCCL gather after the kernel")}
    if peer is not None:
        dev_res = None
        peer.close()
    del part, T
    free_pins(x)
    torch.cuda.empty_cache()
    return line


def run_build_config(x, args, cfg, steps, warmup, do_cpu=True):
    """C5: Voxel_Grid(Model, Domain) cell lists on the GPU vs the CPU hierarchical constructor (rank 0 only)."""
    if x.rank != 0:
        return None
    torch, L, hb = x.torch, x.L, x.hb
    mesh = get_mesh(cfg)
    dom = cfg["args"][0]
    T = hb.Topology.from_mesh(mesh)
    g = None
    for _ in range(max(1, warmup)):
        g = hb.Voxel_Grid([T], dom)
    torch.cuda.synchronize()
    l0 = hb.launch_count()
    walls = []
    for _ in range(steps):
        del g
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = hb.Voxel_Grid([T], dom)
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - t0)
    launches = hb.launch_count() - l0
    wall = float(np.median(walls))
    kms = g.build_ms()[0] or None
    _, _, ct, K = g.info()
    ncells = int(ct[0]) * int(ct[1]) * int(ct[2])
    B = 2 * mesh.P * 128 + 12 * ncells + 4 * K
    peak, peak_kind = peak_hbm()
    cpu_line, parity = None, None
    if do_cpu:
        from oracle import hare_oracle as ho
        cores = os.cpu_count() or 1
        To = ho.Topology.from_mesh(mesh)
        md = int(round(np.log2(dom)))
        t0 = time.perf_counter()
        og = ho.Voxel_Grid(To, md, mode="hier", avg_polys=0, nthreads=cores)
        dt = time.perf_counter() - t0
        off, pol = g.csr(); ooff, opol = og.csr()
        assert np.array_equal(off, ooff) and np.array_equal(pol, opol), "C5: GPU cell lists differ from the oracle's hierarchical build"
        parity = {"csr_equal": True, "cells": ncells, "pairs": int(K)}
        cpu_line = {"value": mesh.P / dt / 1e6, "unit": "Mpolys/s", "cores": cores, "kind": "port",
                    "sample": f"one full hierarchical build Voxel_Grid(Model, MaxDomain={md}, Avg_polys=0) of all {mesh.P} polygons, {dt:.1f} s"}
    t_roof = (kms if kms else wall * 1e3) * 1e-3
    line = {"metric": "Mpolys/s Voxel_Grid build", "value": mesh.P / wall / 1e6, "unit": "Mpolys/s", "n_gpus": 1, "steps": steps, "warmup": warmup,
            "ms_per_step": wall * 1e3, "kernel_ms": kms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, mesh, 1, {"cells": ncells, "pairs": int(K)}),
            "roofline": {"bound": "hbm", "achieved": B / t_roof / 1e9, "peak": peak, "unit": "GB/s", "frac": B / t_roof / 1e9 / peak, "traffic": None,
                         "peak_kind": peak_kind, "kernel": "vg_bin_kernel (count + scatter), scan, vg_finish_cells", "bytes_per_build": B,
                         "formula": "2*P*128 + 12*Ncells + 4*K (SURVEY.md 8(d))", "timed": "kernels (CUDA events inside the build)" if kms else "host wall time"},
            "cpu_baseline": cpu_line, "e2e": {"value": mesh.P / wall / 1e6, "unit": "Mpolys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
                                               "note": "hare_voxelgrid_build wall time; the Topology is already on the device, only the pair count returns"},
            "gpu_launches": int(launches), "parity": parity, "wall_ms_all": [w * 1e3 for w in walls]}
    del g, T
    return line


def run_c1(x, args, cfg):
    """C1 parity configuration + the boundary's cliffs: single-ray Shoot latency, pageable vs page-locked batches."""
    line = run_shoot_config(x, args, cfg, steps=5, warmup=3, cpu_seconds=3.0, do_e2e=True, do_cpu=True)
    if x.rank != 0:
        return None
    hb = x.hb
    from hare_b200.harness import rays_from_sources
    mesh = get_mesh(cfg)
    T = hb.Topology.from_mesh(mesh)
    g = hb.Voxel_Grid([T], 10)
    o, d = rays_from_sources(2000, get_sources(cfg), stream=1)
    R = [hb.Ray(*o[i], *d[i], Ray_ID=i + 1) for i in range(len(o))]
    for r in R[:200]:
        g.Shoot(r, 0)
    t0 = time.perf_counter()
    for r in R:
        g.Shoot(r, 0)
    single_us = (time.perf_counter() - t0) / len(R) * 1e6
    # pageable vs page-locked host arrays through hare_shoot_batch, 4 M rays
    n = 4_000_000
    op, dp = rays_from_sources(n, get_sources(cfg), stream=1)
    t0 = time.perf_counter(); g.Shoot_Batch(op, dp); g.Shoot_Batch(op, dp); pageable = 2 * n / (time.perf_counter() - t0) / 1e6
    line["boundary"] = {"single_ray_Shoot_us": single_us, "single_ray_Shoot_per_s": 1e6 / single_us,
                        "note": "Spatial_Partition.Shoot(Ray) through the C ABI is a one-ray batch: H2D + launch + D2H + sync per call; use the batched overload",
                        "shoot_batch_pageable_Mrays_s": pageable, "shoot_batch_pinned_Mrays_s": line["e2e"]["value"] if line.get("e2e") else None}
    return line


def run_b200(args):
    x = make_ctx(args)
    cfg = cfg_of(args)
    do_cpu = not args.no_cpu_baseline
    if cfg["kind"] == "build":
        line = run_build_config(x, args, cfg, args.steps, args.warmup, do_cpu)
    elif cfg["name"] == "C1":
        line = run_c1(x, args, cfg)
    else:
        line = run_shoot_config(x, args, cfg, args.steps, args.warmup, args.cpu_seconds, do_e2e=not args.no_e2e, do_cpu=do_cpu, gather=args.gather)
    if x.world == 1 and not args.no_extras and args.config == "C3" and line is not None:
        # the other BASELINE configurations, short, so that every round's driver run has them side by side
        others = {}
        for name, st, wu, cs in (("C2", 3, 3, 4.0), ("C4vg", 3, 3, 4.0), ("C4kd", 3, 3, 4.0)):
            c2 = cfg_of(args, name)
            try:
                others[name] = run_shoot_config(x, args, c2, st, wu, cs, do_e2e=(name == "C2"), do_cpu=do_cpu)
            except AssertionError:
                raise
            except Exception as e:   # e.g. out of host memory on a small box: say so instead of hiding the main line
                others[name] = {"error": repr(e)}
        try:
            others["C5"] = run_build_config(x, args, cfg_of(args, "C5"), 5, 2, do_cpu)
        except AssertionError:
            raise
        except Exception as e:
            others["C5"] = {"error": repr(e)}
        try:
            others["C1"] = run_c1(x, args, cfg_of(args, "C1"))
        except AssertionError:
            raise
        except Exception as e:
            others["C1"] = {"error": repr(e)}
        line["other_configs"] = others
    if x.rank == 0 and line is not None:
        print(json.dumps(line))
    if x.world > 1:
        x.dist.barrier()
        x.dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
