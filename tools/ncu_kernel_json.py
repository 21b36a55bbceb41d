#!/usr/bin/env python
"""Turn one kernel of an `ncu --set full` report into the small JSON bench.py reads its `roofline.traffic` from.

    python tools/ncu_kernel_json.py gpurun_out/prof.ncu-rep profiles/r1_ncu_bench_kernel.json --mesh 50k --domain 64 --rays 10000000 --order 50 \
        --command "ncu --set full ... python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
"""
import argparse
import csv
import json
import subprocess

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("out")
    ap.add_argument("--mesh", default="50k"); ap.add_argument("--domain", type=int, default=64)
    ap.add_argument("--rays", type=int, default=10_000_000); ap.add_argument("--order", type=int, default=50)
    ap.add_argument("--command", default="")
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
    by = lambda k: float(j[k]["value"].replace(",", "")) * SCALE[j[k]["unit"]]
    ms = float(j["gpu__time_duration.sum"]["value"].replace(",", ""))
    j["traffic_bytes_per_launch"] = by("dram__bytes_read.sum") + by("dram__bytes_write.sum")
    j["l2_bytes_per_launch"] = float(j["lts__t_sectors.sum"]["value"].replace(",", "")) * 32
    j["l2_GBps"] = j["l2_bytes_per_launch"] / (ms * 1e-3) / 1e9
    j["command"] = a.command
    j["config"] = {"mesh": a.mesh, "domain": a.domain, "rays": a.rays, "order": a.order}
    json.dump(j, open(a.out, "w"), indent=1)
    print(json.dumps({k: j[k] for k in ("traffic_bytes_per_launch", "l2_GBps")}))


if __name__ == "__main__":
    main()
