#!/usr/bin/env python
"""Round-2 summaries from the gpurun_out/ captures of profiles/capture_r02.sh (tracked output: profiles/r02_*.md).

    python profiles/summarize_r02.py [tag]

* <tag>_launches.md / <tag>_launches_c4.md: every kernel of the step with its launches, average duration, share of the step,
  DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) and the achieved DRAM rate against the measured HBM
  peak (MEASURED_PEAKS.json) - the per-kernel table VERDICT r01 asked for. ncu serialises launches and runs them cold:
  compare SHARES, not absolute times, with the bench line.
* <tag>_<kernel>.md: key metrics of the `--set full` capture of that kernel (raw page exported on the GPU box).
"""
import collections
import csv
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "gpurun_out")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]

# why a kernel sits below 0.6 of the HBM roofline (SURVEY.md 8d asks for a reason per memory-bound kernel)
WHY = {
    "k_gs_exact": "conversion-pipe bound (27 F2F.F64.F32 + 12 F2F.F32.F64 per row and lane, ~3 lanes/clk/SMSP) and one grid barrier per colour phase; DESIGN.md 4.3",
    "k_np_hull_warp<1, 0>": "FP64 instruction bound: 27 pruned SAT axes x (8 + 6) vertex projections per box-pillar task; 136 MB of DRAM for 0.44 ms of arithmetic",
    "k_np_hull_warp<1, 1>": "clipping of the queued pillar tasks: dependent shared-memory passes per task, latency bound",
    "k_np_hull_warp<0, 0>": "FP64 instruction bound SAT of hull-hull tasks (15 axes box-box)",
    "k_np_hull_warp<0, 1>": "clipping of the queued hull-hull tasks, latency bound",
    "k_np_tasks": "walks the heightfield cells under every pair AABB: data-dependent loops, L2-resident inputs (the pairs and 66k pillar records)",
    "k_rows_build": "one thread per unit assembles its rows in f64 (SPOOK terms, 4 cross products, two I^-1 r products per row): FP64 + 206 registers, writes 126 MB",
    "k_schedule": "Jones-Plassmann colouring: ~10 rounds x 2 grid barriers over 215k units, latency of the barrier chain, not bandwidth",
    "k_bp_small": "hash-grid neighbour walk: 27 cells x a few bodies per body, L2-resident body records, divergent sphere tests",
    "k_np_finalize": "per-contact material lookup + friction parameters: scattered 16-byte gathers by body index",
    "k_np_sphere_hull": "sphere-cylinder resolver in f64, one thread per task",
    "k_np_sphere_pillar": "sphere-pillar resolver in f64, one thread per task",
    "k_np_sphere_box": "sphere-box resolver in f64, one thread per task",
    "k_units_build": "one pass over the tasks; small",
    "k_scan_onepass": "single-pass look-back scan: tiles wait for their predecessors, ~20 us fixed latency per scan whatever the size",
    "k_integrate": "the one streaming kernel of the step: 100k bodies x ~250 B; too small to reach steady state (13 us)",
    "k_prestep": "streaming, too small to reach steady state",
    "k_bp_world_all": "all-pairs Naive broadphase of a small world, one warp per body with ballots: latency of one pass over <= 64 bodies",
    "k_schedule_worlds": "Jones-Plassmann colouring per world (one warp, claims in shared memory, units in registers): a few dependent rounds, no grid barrier",
    "k_world_keys": "per-world min / max of the contact keys (world-local colouring priorities): one atomic per warp and world",
    "k_gs_world_exact": "one warp per world, rows in shared memory: conversion-pipe latency of a lone warp per SM sub-partition (DRAM is only touched when the rows are staged)",
}


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([A-Za-z0-9_]+(<[^>]*>)?)", name)
    return m.group(1) if m else name


def launches(tag, path, out_name, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()  # (id) -> dict
    for row in csv.DictReader(lines):
        d = per.setdefault(row["ID"], {"kernel": short(row["Kernel Name"])})
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        u = row["Metric Unit"]
        if row["Metric Name"] == "gpu__time_duration.sum":
            v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[u]
        elif row["Metric Name"].startswith("dram__bytes"):
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        d[row["Metric Name"]] = v
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["kernel"], {"n": 0, "us": 0.0, "bytes": 0.0, "occ": 0.0})
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a["occ"] += d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
    T = sum(a["us"] for a in agg.values())
    out = [f"# {tag}: {title}\n",
           f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, {len(per)} consecutive launches "
           f"(cold-cache, serialised: compare SHARES). Achieved = DRAM bytes / duration; peak = {PEAK:.1f} GB/s (MEASURED_PEAKS.json).\n",
           "| kernel | launches | avg us | share | DRAM MB / launch | achieved GB/s | frac of HBM peak | warps active % | below 0.6 because |", "|---|---|---|---|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
        gbs = a["bytes"] / (a["us"] * 1e-6) / 1e9 if a["us"] > 0 else 0.0
        out.append(f"| `{k}` | {a['n']} | {a['us'] / a['n']:.1f} | {100 * a['us'] / T:.1f} % | {a['bytes'] / a['n'] / 1e6:.2f} | {gbs:.0f} | {gbs / PEAK:.3f} | "
                   f"{a['occ'] / a['n']:.0f} | {WHY.get(k, '')} |")
    out.append(f"\ntotal {T:.0f} us over {len(per)} launches")
    open(os.path.join(HERE, out_name), "w").write("\n".join(out) + "\n")
    return agg


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def report(tag, name, path):
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[-1]
    kernel = short(vals[hdr.index("Kernel Name")])
    out = [f"# {tag}: ncu --set full --clock-control none, kernel `{kernel}` (launch 10 of its name in the settled-pile bench)\n", "| metric | value | unit |", "|---|---|---|"]
    got = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            out.append(f"| {k} | {vals[i]} | {units[i]} |")
            got[k] = (vals[i], units[i])
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    try:
        b = sum(float(got[k][0].replace(",", "")) * scale[got[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t = float(got["gpu__time_duration.sum"][0].replace(",", "")) * scale[got["gpu__time_duration.sum"][1]]
        out.append(f"\nDRAM traffic {b / 1e6:.1f} MB in {t * 1e6:.1f} us = {b / t / 1e9:.0f} GB/s = {b / t / 1e9 / PEAK:.3f} of the measured HBM peak ({PEAK:.1f} GB/s).")
        res = {"kernel": kernel, "dram_bytes": b, "seconds": t}
    except Exception:
        res = None
    if kernel in WHY:
        out.append(f"\nWhat bounds it: {WHY[kernel]}.")
    open(os.path.join(HERE, f"{tag}_{name}.md"), "w").write("\n".join(out) + "\n")
    return res


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    launches(tag, os.path.join(OUT, f"{tag}_launches.csv"), f"{tag}_launches.md",
             "per-kernel device time and DRAM traffic, settled 100k pile of config 3 (~4 steps of the timed region)")
    p = os.path.join(OUT, f"{tag}_launches_c4.csv")
    if os.path.exists(p):
        launches(tag, p, f"{tag}_launches_c4.md", "per-kernel device time and DRAM traffic, config 4, one 512-world shard (~2 steps)")
    traffic = {}
    for f in sorted(os.listdir(OUT)):
        m = re.match(rf"{tag}_(k_[a-z_]+)\.raw\.csv$", f)
        if m:
            r = report(tag, m.group(1), os.path.join(OUT, f))
            if r:
                traffic[r["kernel"]] = r
    json.dump(traffic, open(os.path.join(HERE, f"{tag}_ncu_traffic.json"), "w"), indent=1)
    print("wrote", [f for f in sorted(os.listdir(HERE)) if f.startswith(tag)])
