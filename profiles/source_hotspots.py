"""Per-source-line hot spots of one kernel launch from an .ncu-rep captured with --import-source on.

    python profiles/source_hotspots.py gpurun_out/full5.ncu-rep k_shade [launch_index] [top]
"""
import csv
import subprocess
import sys


def main(rep, kernel, skip=0, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kernel}",
                          "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
    fname, hdr, lines = None, None, []
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-":
            g = lambda k: float(r[hdr.index(k)] or 0)
            lines.append((fname, int(r[0]), r[1].strip()[:90], g("# Samples"), g("Instructions Executed"), g("Thread Instructions Executed"),
                          g("stall_long_sb"), g("stall_wait"), g("stall_no_inst"), g("stall_branch_resolving"), g("stall_short_sb"), g("stall_lg")))
    ts = sum(l[3] for l in lines); ti = sum(l[4] for l in lines); tt = sum(l[5] for l in lines)
    print(f"{kernel} launch {skip}: samples {ts:.0f}, warp instr {ti:.0f}, threads/instr {tt / max(ti, 1):.2f}")
    byfile = {}
    for l in lines:
        a = byfile.setdefault(l[0], [0, 0, 0]); a[0] += l[3]; a[1] += l[4]; a[2] += l[5]
    for f, a in byfile.items():
        print(f"  {f:24s} samples {100 * a[0] / ts:5.1f}%  instr {100 * a[1] / ti:5.1f}%  thr/instr {a[2] / max(a[1], 1):5.1f}")
    print("  file:line                 samp%  inst%  thr/i  longsb wait noinst branch shortsb lg | source")
    for l in sorted(lines, key=lambda l: -l[3])[:top]:
        print(f"  {l[0][5:-4]:>10s}:{l[1]:<5d} {100 * l[3] / ts:5.1f} {100 * l[4] / ti:5.1f} {l[5] / max(l[4], 1):5.1f}   "
              f"{l[6]:5.0f} {l[7]:5.0f} {l[8]:5.0f} {l[9]:5.0f} {l[10]:5.0f} {l[11]:5.0f} | {l[2]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0, int(sys.argv[4]) if len(sys.argv) > 4 else 40)
