"""DRAM traffic per ray of the trace kernels from an `ncu --set full` report.

    python profiles/traffic_from_ncu.py gpurun_out/full.ncu-rep profiles/traffic_r1.json

ncu prints byte counters with a per-column unit (byte / Kbyte / Mbyte / Gbyte) that differs from
launch to launch and column to column: every value is scaled by its own unit here.  Rays per
launch = grid size x block size (an upper bound: the last block is partly empty).
"""
import csv
import json
import subprocess
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
                                     "gpu__time_duration.sum")}
    res = {}
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
        rd = float(r[col["dram__bytes_read.sum"]]) * SCALE[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * SCALE[units[col["dram__bytes_write.sum"]]]
        n = int(r[col["launch__grid_size"]]) * int(r[col["launch__block_size"]])
        e = res.setdefault(name, {"launches": 0, "threads": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        e["launches"] += 1; e["threads"] += n; e["dram_read_bytes"] += rd; e["dram_write_bytes"] += wr
    for e in res.values():
        e["dram_bytes_per_thread"] = (e["dram_read_bytes"] + e["dram_write_bytes"]) / max(e["threads"], 1)
    doc = {"source": rep.split("/")[-1] + " (ncu --set full, per-column units applied)", "kernels": res,
           "k_trace_closest_dram_bytes_per_ray": res.get("k_trace_closest", {}).get("dram_bytes_per_thread"),
           "note": "dram__bytes_read.sum + dram__bytes_write.sum over the captured launches / (grid x block) threads in them; bench.py scales the "
                   "k_trace_closest figure by the rays of its average launch"}
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    for k, e in res.items():
        print(f"{k:18s} launches {e['launches']:2d}  DRAM {e['dram_bytes_per_thread']:8.1f} B/thread")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
