#!/usr/bin/env python
"""bench.py -- Mrays/s of closest-hit Shoot on BASELINE.json's configs[1] (C2):
procedural auditorium (~50k polygons), Voxel_Grid, 10M rays x 50-order specular chains.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference

One "step" = one pass of the hot path over the whole ray batch: every ray's chain of up to
`order` Shoots (each bounce is one Shoot).  `value` is device-resident
throughput (rays already in HBM), `e2e` goes through the host-buffer C-ABI call
(hare_reflect_chain) with the H2D / D2H copies inside the timed region.

For N > 1 (torchrun, one rank per GPU) the geometry is replicated, every rank owns its own
batch of `--rays` chains (weak scaling) and the per-chain results are gathered to rank 0 over
NCCL inside the timed region (the path's only cross-GPU step).
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=10_000_000, help="chains per GPU per step (C2: 10M)")
    ap.add_argument("--order", type=int, default=50)
    ap.add_argument("--mesh", default="50k")
    ap.add_argument("--domain", type=int, default=64, help="Voxel_Grid Domain (C2 sweeps 32/64/96)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="profiling runs: skip the CPU leg")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the host-buffer leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def ncu_traffic(args):
    """dram__bytes_read.sum + dram__bytes_write.sum of the traversal kernel, per launch, from the committed
    `ncu --set full` capture of this same configuration (profiles/); None for any other configuration."""
    p = os.path.join(ROOT, "profiles", "r1_ncu_bench_kernel.json")
    try:
        j = json.load(open(p))
        c = j["config"]
        if (c["mesh"], c["domain"], c["rays"], c["order"]) == (args.mesh, args.domain, args.rays, args.order):
            return float(j["traffic_bytes_per_launch"])
    except Exception:
        pass
    return None


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return PEAKS_FALLBACK_HBM, "fallback"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
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


def make_workload(args, rank):
    from hare_b200.harness import meshes, rays_from_sources
    mesh = meshes.hall(args.mesh)
    # rank r shoots its own block of the global ray sequence (streams 2..5 = the four sources of C2)
    o, d = rays_from_sources(args.rays, meshes.sources(4), stream=2, first=rank * args.rays)
    return mesh, o, d


def algorithmic_bytes(counters, shots, kind="Voxel_Grid"):
    """B = 56 + E + h*C + 4*L + 128*T per Shoot, C/L/T counted by the CPU oracle."""
    E, h = (36, 8) if kind == "Voxel_Grid" else (52, 64)
    C_, L_, T_ = (float(counters[k]) / shots for k in range(3))
    return 56 + E + h * C_ + 4 * L_ + 128 * T_, dict(cells=C_, entries=L_, tests=T_)


def cpu_leg(args, mesh, o, d, seconds, nthreads):
    """Time the oracle (C++ restatement of the reference's CPU path) on a bounded sample."""
    from oracle import hare_oracle as ho
    To = ho.Topology.from_mesh(mesh)
    g = ho.Voxel_Grid(To, args.domain, mode="fast")
    n0 = min(len(o), 4000 * nthreads)
    t0 = time.perf_counter(); r = g.reflect_chain(o[:n0], d[:n0], args.order, events=False, nthreads=nthreads); dt = time.perf_counter() - t0
    rate = r["nshots"].sum() / dt
    n = int(min(len(o), max(n0, seconds * rate / args.order)))
    t0 = time.perf_counter(); r = g.reflect_chain(o[:n], d[:n], args.order, events=False, nthreads=nthreads); dt = time.perf_counter() - t0
    shots = int(r["nshots"].sum())
    return dict(mrays=shots / dt / 1e6, shots=shots, n=n, seconds=dt, counters=r["counters"], part=g)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    mesh, o, d = make_workload(args, 0)
    per_step = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    from oracle import hare_oracle as ho
    To = ho.Topology.from_mesh(mesh)
    g = ho.Voxel_Grid(To, args.domain, mode="fast")
    n0 = min(len(o), 4000 * cores)
    t0 = time.perf_counter(); r = g.reflect_chain(o[:n0], d[:n0], args.order, events=False, nthreads=cores); dt = time.perf_counter() - t0
    n = int(min(len(o), max(n0, per_step * (r["nshots"].sum() / dt) / args.order)))
    for _ in range(args.warmup):
        g.reflect_chain(o[:n], d[:n], args.order, events=False, nthreads=cores)
    t0 = time.perf_counter(); shots = 0
    for _ in range(args.steps):
        shots += int(g.reflect_chain(o[:n], d[:n], args.order, events=False, nthreads=cores)["nshots"].sum())
    dt = time.perf_counter() - t0
    val = shots / dt / 1e6
    sample = f"{n} chains x {args.order} Shoots per step ({shots // max(1, args.steps)} Shoots/step) of the {args.rays}-chain workload"
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s closest-hit Shoot", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, mesh),
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C++ restatement of Hare's Voxel_Grid.Shoot CPU path (oracle/), std::thread over rays; no .NET runtime exists on this box",
    }))


def workload_config(args, mesh):
    return {"workload": f"C2: procedural auditorium hall-{args.mesh} ({mesh.P} polygons), Voxel_Grid Domain {args.domain}, "
                        f"{args.rays} rays x {args.order}-order specular chains per GPU",
            "partition": "Voxel_Grid", "domain": args.domain, "polygons": mesh.P, "rays_per_gpu": args.rays, "order": args.order,
            "l2": "ray inputs (48 B/ray) exceed L2 each step; geometry (polygons + cells) is L2-resident by design"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import hare_b200 as hb
    from hare_b200._lib import check, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hb.init([local])
    L = lib()

    mesh, o, d = make_workload(args, rank)
    N, order = args.rays, args.order
    cpu_res = None
    if rank == 0 and not args.no_cpu_baseline:
        # the CPU leg runs first, on an otherwise idle host (before pinned buffers and GPU work exist)
        cpu_res = cpu_leg(args, mesh, o, d, args.cpu_seconds, os.cpu_count() or 1)
    if world > 1:
        dist.barrier()
    T = hb.Topology.from_mesh(mesh)
    part = hb.Voxel_Grid([T], args.domain)

    # ---------------- device-resident leg: rays already in HBM --------------------------------------
    o_d = torch.from_numpy(o).to(dev); d_d = torch.from_numpy(d).to(dev)
    fin_o = torch.empty_like(o_d); fin_d = torch.empty_like(d_d)
    nshots = torch.empty(N, dtype=torch.int32, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    if world > 1:
        g_fin_o = torch.empty((world * N, 3), dtype=torch.float64, device=dev) if rank == 0 else None
        g_fin_d = torch.empty((world * N, 3), dtype=torch.float64, device=dev) if rank == 0 else None
        g_ns = torch.empty(world * N, dtype=torch.int32, device=dev) if rank == 0 else None

    def gather():   # X_Event rows of every rank -> rank 0, in rank order (hare_b200/dist.py; same helper as the gloo test)
        dist.gather(fin_o, list(g_fin_o.chunk(world)) if rank == 0 else None, dst=0)
        dist.gather(fin_d, list(g_fin_d.chunk(world)) if rank == 0 else None, dst=0)
        dist.gather(nshots, list(g_ns.chunk(world)) if rank == 0 else None, dst=0)

    def cur_stream():
        # torch's default stream is the legacy default stream (handle 0); the C ABI treats NULL as "the
        # partition's own stream", so name the legacy stream explicitly (cudaStreamLegacy == 0x1).
        s = torch.cuda.current_stream().cuda_stream
        return s if s else 1

    def step_device():
        stream = cur_stream()
        check(L.hare_reflect_chain_device(part._h, o_d.data_ptr(), d_d.data_ptr(), N, order, None, None,
                                          fin_o.data_ptr(), fin_d.data_ptr(), nshots.data_ptr(), total.data_ptr(), None, C.c_void_p(stream)),
              "hare_reflect_chain_device")
        if world > 1:   # gather of per-chain results to rank 0 over NVLink (NCCL)
            gather()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    sync_all()
    total.zero_()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = hb.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    ev[0].record()
    for k in range(args.steps):
        # inner events bracket the traversal kernel alone (for the roofline line); outer ones the step
        kev[k][0].record()
        stream = cur_stream()
        check(L.hare_reflect_chain_device(part._h, o_d.data_ptr(), d_d.data_ptr(), N, order, None, None,
                                          fin_o.data_ptr(), fin_d.data_ptr(), nshots.data_ptr(), total.data_ptr(), None, C.c_void_p(stream)),
              "hare_reflect_chain_device")
        kev[k][1].record()
        if world > 1:
            gather()
        ev[k + 1].record()
    sync_all()
    launches = hb.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms_total = ev[0].elapsed_time(ev[-1])
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    shots_rank = int(total.item())
    tms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    sh = torch.tensor([shots_rank], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(sh, op=dist.ReduceOp.SUM)
    ms_total = float(tms.item()); shots_all = int(sh.item())
    value = shots_all / (ms_total * 1e-3) / 1e6

    # ---------------- end-to-end leg: host buffers through the public C-ABI call --------------------------
    e2e = None
    if not args.no_e2e:
        o_h = torch.from_numpy(o).pin_memory(); d_h = torch.from_numpy(d).pin_memory()
        fo_h = torch.empty((N, 3), dtype=torch.float64).pin_memory(); fd_h = torch.empty((N, 3), dtype=torch.float64).pin_memory()
        ns_h = torch.empty(N, dtype=torch.int32).pin_memory()
        tot = C.c_uint64()

        def step_host():
            check(L.hare_reflect_chain(part._h, o_h.data_ptr(), d_h.data_ptr(), N, order, None, None,
                                       fo_h.data_ptr(), fd_h.data_ptr(), ns_h.data_ptr(), C.byref(tot), None), "hare_reflect_chain")
            return tot.value
        for _ in range(min(args.warmup, 2)):
            step_host()
        sync_all()
        e_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter(); e_shots = 0
        for _ in range(e_steps):
            e_shots += step_host()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev); es = torch.tensor([e_shots], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX); dist.all_reduce(es, op=dist.ReduceOp.SUM)
        e2e = {"value": int(es.item()) / float(tt.item()) / 1e6, "unit": "Mrays/s",
               "h2d_bytes_per_step": int(world * N * 48), "d2h_bytes_per_step": int(world * N * 52),
               "steps": e_steps, "api": "hare_reflect_chain (host buffers, pinned)"}
        assert np.array_equal(ns_h.numpy(), nshots.cpu().numpy()), "host-buffer and device-resident legs disagree"

    if rank == 0:
        peak, peak_kind = peak_hbm()
        cpu = None
        bytes_per_shoot, avg = None, None
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            c = cpu_res
            bytes_per_shoot, avg = algorithmic_bytes(c["counters"], c["shots"])
            cpu = {"value": c["mrays"], "unit": "Mrays/s", "cores": cores, "kind": "port",
                   "sample": f"first {c['n']} chains x {order} Shoots ({c['shots']} Shoots, {c['seconds']:.1f} s) of the {N}-chain workload"}
            # parity spot-check of the timed device leg against the same oracle run (not timed)
            ref = c["part"].reflect_chain(o[:20000], d[:20000], order, events=False, nthreads=cores)
            assert np.array_equal(ref["nshots"], nshots[:20000].cpu().numpy()) and np.array_equal(ref["o"], fin_o[:20000].cpu().numpy()), \
                "bench result differs from the oracle"
        else:
            # counters from the GPU's own walk (cells identical to the oracle's; entries/tests are upper bounds of the mailboxed reference)
            r = part.Reflect_Chain(o[:200000], d[:200000], order, events=False, counters=True)
            bytes_per_shoot, avg = algorithmic_bytes(r["counters"], r["total_shots"])
        shots_per_launch = shots_rank / args.steps
        achieved = shots_per_launch * bytes_per_shoot / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": "Mrays/s closest-hit Shoot", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, mesh),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args),
                         "peak_kind": peak_kind, "kernel": ("vg_walk_kernel<CHAIN>" if os.environ.get("HARE_VG_WAVE") == "0" else "vg_wave_kernel<CHAIN>"), "kernel_ms": kernel_ms,
                         "bytes_per_shoot": bytes_per_shoot, "per_shoot": avg,
                         "formula": "56 + 36 + 8*cells + 4*entries + 128*tests (SURVEY.md 8(d)), oracle-counted"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "shots_per_step": shots_all // args.steps,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
