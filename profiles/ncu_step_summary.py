"""Summarise an ncu metrics capture of ONE bench step (tests/ncu_target.py 3 0 16) per kernel.

Capture (on the GPU box, after the same command has exited 0 without ncu):

    ncu --clock-control none --csv --log-file gpurun_out/ncu_step.csv --metrics <METRICS below> \
        python tests/ncu_target.py 3 0 16

    python profiles/ncu_step_summary.py gpurun_out/ncu_step.csv gpurun_out/ncu_target_stats.json profiles/ncu_r2_step.json

Output: per kernel class the launches, summed device time, warp instructions, thread instructions, DRAM
bytes, L2 bytes; per ray of that class (closest-hit kernel: path + MIS rays; any-hit: shadow rays; shade
kernels: per path vertex = closest-hit path ray) the instruction / byte figures bench.py combines with its
own live ray counts and kernel times (roofline_issue, l2, roofline.traffic)."""
import csv
import json
import re
import sys
from collections import defaultdict

METRICS = ("gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "lts__t_bytes.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,"
           "sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,"
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio")


def to_float(v, unit):
    v = float(v.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
    return v * scale.get(unit, 1.0)


def main(csv_path, stats_path, out_path):
    rows = []
    with open(csv_path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        rows.append(r)
    per_launch = defaultdict(dict)
    names = {}
    for r in rows:
        lid = int(r["ID"])
        names[lid] = re.sub(r"<.*", "", r["Kernel Name"].split("(")[0]).replace("void ", "").strip()
        per_launch[lid][r["Metric Name"]] = to_float(r["Metric Value"], r["Metric Unit"])
    agg = defaultdict(lambda: defaultdict(float))
    for lid, m in per_launch.items():
        k = names[lid]
        a = agg[k]
        t = m.get("gpu__time_duration.sum", 0.0)
        a["launches"] += 1
        a["ms"] += t
        a["warp_inst"] += m.get("smsp__inst_executed.sum", 0.0)
        a["thread_inst"] += m.get("smsp__thread_inst_executed.sum", 0.0)
        a["dram_bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        a["l2_bytes"] += m.get("lts__t_bytes.sum", 0.0)
        for key, name in (("l1_hit", "l1tex__t_sector_hit_rate.pct"), ("l2_hit", "lts__t_sector_hit_rate.pct"),
                          ("issue_active", "smsp__issue_active.avg.pct_of_peak_sustained_elapsed"), ("warps_active", "sm__warps_active.avg.pct_of_peak_sustained_active"),
                          ("fma_pipe", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), ("alu_pipe", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active")):
            a[key + "_x_ms"] += m.get(name, 0.0) * t          # time-weighted
    stats = json.load(open(stats_path))
    # the ray counts must be those of the CAPTURED run: ncu_target.py writes them during the capture, and any later run of
    # it overwrites the file (round 2 shipped a summary whose counts came from a 4-spp run beside a 16-spp capture)
    for kname, key in (("k_trace_closest", "launches_closest"), ("k_trace_any", "launches_any")):
        seen = int(agg[kname]["launches"]) if kname in agg else 0
        if seen != int(stats[key]):
            sys.exit(f"{stats_path} does not belong to {csv_path}: {kname} launched {seen} times in the capture, {stats[key]} in the stats file")
    units = {"k_trace_closest": stats["rays_closest_kernel"], "k_trace_any": stats["rays_any_kernel"], "k_shade_a": stats["rays_closest_kernel"],
             "k_shade_b": stats["rays_closest_kernel"]}
    total_ms = sum(a["ms"] for a in agg.values())
    kernels = {}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        e = {"launches": int(a["launches"]), "ms_under_ncu": round(a["ms"], 3), "share_of_step": round(a["ms"] / total_ms, 4),
             "warp_inst": a["warp_inst"], "threads_per_inst": round(a["thread_inst"] / max(a["warp_inst"], 1), 2),
             "dram_bytes": a["dram_bytes"], "l2_bytes": a["l2_bytes"]}
        for key in ("l1_hit", "l2_hit", "issue_active", "warps_active", "fma_pipe", "alu_pipe"):
            e[key + "_pct"] = round(a[key + "_x_ms"] / max(a["ms"], 1e-9), 2)
        if k in units and units[k]:
            n = units[k]
            e.update(units=n, warp_inst_per_ray=round(a["warp_inst"] / n, 2), dram_bytes_per_ray=round(a["dram_bytes"] / n, 2), l2_bytes_per_ray=round(a["l2_bytes"] / n, 2))
        kernels[k] = e
    out = {"source": "ncu metrics pass of `python tests/ncu_target.py %d %d %d` (one bench step), profiles/ncu_r2_step.json; per-launch times under ncu are cold-cache "
                     "and serialised: shares, not absolutes" % (stats["config"], stats["level"], stats["spp"]),
           "metrics": METRICS, "workload": stats, "kernels": kernels}
    json.dump(out, open(out_path, "w"), indent=1)
    for k, e in kernels.items():
        print(f"{k:22s} launches {e['launches']:4d}  {e['ms_under_ncu']:9.3f} ms  share {e['share_of_step']:.3f}  thr/inst {e['threads_per_inst']:5.2f}  "
              f"issue {e['issue_active_pct']:5.1f}%  L1 {e['l1_hit_pct']:5.1f}%  L2 {e['l2_hit_pct']:5.1f}%  dram {e['dram_bytes'] / 1e6:9.1f} MB  l2 {e['l2_bytes'] / 1e6:9.1f} MB")


if __name__ == "__main__":
    if len(sys.argv) == 2 and sys.argv[1] == "--metrics":
        print(METRICS)
    else:
        main(*sys.argv[1:4])
