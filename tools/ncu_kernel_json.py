#!/usr/bin/env python
"""Turn one kernel of an `ncu --set full` report into the small JSON bench.py reads `roofline.traffic` / `roofline.actual` from.

    python tools/ncu_kernel_json.py gpurun_out/r2_c3.ncu-rep profiles/r2_ncu_C3.json --config C3 --mesh 500k --part Octree --args 7 32 \
        --rays 4000000 --command "ncu --set full ... python bench.py --config C3 --rays 4000000 ..."

`traffic_bytes_per_launch` is the capture's DRAM bytes; bench.py scales it to its own launch by the ray count
(`traffic_bytes_per_ray`).  `bound` names the limiter the counters point at: "hbm" when DRAM throughput exceeds 60 % of peak,
"l2" when the L2 does, otherwise "issue" (instruction issue at partial SIMT utilisation, waiting on long-scoreboard loads).
"""
import argparse
import csv
import json
import subprocess

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("out")
    ap.add_argument("--config", default="C3"); ap.add_argument("--mesh", default="500k"); ap.add_argument("--part", default="Octree")
    ap.add_argument("--args", type=int, nargs="+", default=[7, 32]); ap.add_argument("--rays", type=int, default=4_000_000)
    ap.add_argument("--order", type=int, default=1); ap.add_argument("--command", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    j = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP:
            j[h] = {"unit": u, "value": v}
        if h == "Kernel Name":
            j["kernel"] = v
    num = lambda k: float(j[k]["value"].replace(",", ""))
    by = lambda k: num(k) * SCALE[j[k]["unit"]]
    ms = num("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(j["gpu__time_duration.sum"]["unit"], 1.0)
    j["traffic_bytes_per_launch"] = by("dram__bytes_read.sum") + by("dram__bytes_write.sum")
    j["traffic_bytes_per_ray"] = j["traffic_bytes_per_launch"] / (a.rays * a.order)
    j["l2_bytes_per_launch"] = num("lts__t_sectors.sum") * 32
    dram_gbs = j["traffic_bytes_per_launch"] / (ms * 1e-3) / 1e9
    l2_gbs = j["l2_bytes_per_launch"] / (ms * 1e-3) / 1e9
    dram_pct = num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    l2_pct = num("lts__throughput.avg.pct_of_peak_sustained_elapsed")
    j["actual"] = {
        "dram_GBs": dram_gbs, "dram_pct_of_peak": dram_pct, "l2_GBs": l2_gbs, "l2_pct_of_peak": l2_pct,
        "l2_hit_pct": num("lts__t_sector_hit_rate.pct"), "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"),
        "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "lanes_per_inst": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "long_scoreboard_per_issue": num("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        "fp64_pipe_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "warp_inst_per_ray": num("smsp__inst_executed.sum") / (a.rays * a.order),
        "registers": num("launch__registers_per_thread"), "kernel_ms_under_ncu": ms, "rays_in_capture": a.rays * a.order,
    }
    j["bound"] = "hbm" if dram_pct > 60 else ("l2" if l2_pct > 60 else "issue")
    j["command"] = a.command
    j["config"] = {"name": a.config, "mesh": a.mesh, "part": a.part, "args": a.args, "rays": a.rays, "order": a.order}
    json.dump(j, open(a.out, "w"), indent=1)
    print(json.dumps({"bound": j["bound"], **{k: round(v, 3) for k, v in j["actual"].items()}}))


if __name__ == "__main__":
    main()
