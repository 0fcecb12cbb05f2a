#!/usr/bin/env python
"""Turns the gpurun_out/ captures into the tracked summaries under profiles/.

    python profiles/summarize.py <round-tag> <launches.csv> <name=report.ncu-rep> ...

Writes profiles/<tag>_launches.md (per-kernel share of the step from the `gpu__time_duration.sum` pass),
profiles/<tag>_<name>.md (key ncu metrics + top stall instructions) and profiles/ncu_traffic.json.
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def launches(tag, path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    out = [f"# {tag}: per-kernel device time (ncu --metrics gpu__time_duration.sum --clock-control none, {sum(cnt.values())} launches "
           f"of the contact-rich regime; cold-cache / serialised: compare SHARES)\n", "| kernel | launches | avg us | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        out.append(f"| `{k}` | {cnt[k]} | {v / cnt[k]:.1f} | {100 * v / T:.1f} % |")
    out.append(f"\ntotal {T:.0f} us over {sum(cnt.values())} launches")
    open(os.path.join(HERE, f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")
    return tot, cnt


def report(tag, name, path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    kernel = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else name
    out = [f"# {tag}: ncu --set full --clock-control none, kernel `{kernel}`\n", "| metric | value | unit |", "|---|---|---|"]
    metrics = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            out.append(f"| {h} | {v} | {u} |")
            metrics[h] = (v, u)
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 2:
        sh = srows[1]
        idx = {h: i for i, h in enumerate(sh)}
        data = srows[2:]
        tot = sum(int(r[idx["# Samples"]] or 0) for r in data) or 1
        out += ["\n## top stall instructions (warp-state samples)\n", "| share | SASS | stall reasons |", "|---|---|---|"]
        for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:14]:
            st = {k: int(r[idx[k]]) for k in sh if k.startswith("stall_") and "Not Issued" not in k and int(r[idx[k]] or 0) > 0}
            out.append(f"| {100 * int(r[idx['# Samples']]) / tot:.1f} % | `{r[idx['Source']].strip()[:60]}` | {st} |")
    open(os.path.join(HERE, f"{tag}_{name}.md"), "w").write("\n".join(out) + "\n")
    return metrics


def to_bytes(v, u):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag, sys.argv[2])
    traffic = {}
    for spec in sys.argv[3:]:
        name, path = spec.split("=")
        m = report(tag, name, path)
        if "dram__bytes_read.sum" in m:
            traffic[name] = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
    if "k_gs_fast" in traffic:
        json.dump({"k_gs_dram_bytes_per_launch": traffic["k_gs_fast"], "source": f"profiles/{tag}_k_gs_fast.md", "all": traffic},
                  open(os.path.join(HERE, "ncu_traffic.json"), "w"), indent=1)
    print(traffic)
