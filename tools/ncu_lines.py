"""Per-source-line instruction counts of one kernel from an .ncu-rep (captured with --import-source on):
    python tools/ncu_lines.py report.ncu-rep [top_n]
Prints totals, IPC-related headline metrics and the source lines that execute the most instructions."""
import collections
import csv
import io
import subprocess
import sys


def page(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = page(rep, "--page", "raw")
    ix = {h: i for i, h in enumerate(raw[0])}
    for k in ("gpu__time_duration.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
              "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"):
        if k in ix:
            print(f"{k}: {raw[2][ix[k]]} {raw[1][ix[k]]}")
    rows = page(rep, "--page", "source", "--print-source", "cuda,sass")
    cur, hdr, line, src = None, None, None, ""
    agg = collections.defaultdict(lambda: [0, 0, ""])
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr and len(r) == len(hdr):
            if r[0] != "":
                line, src = int(r[0]), r[1]
            try:
                e, s = int(r[iE]), int(r[iS])
            except ValueError:
                continue
            a = agg[(cur, line)]
            a[0] += e
            a[1] += s
            a[2] = src
    tot = sum(v[0] for v in agg.values())
    print("warp instructions attributed:", tot)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0]:10d} {100 * v[0] / tot:5.1f}% smp {v[1]:5d} {k[0]}:{k[1]} {v[2].strip()[:96]}")


if __name__ == "__main__":
    main()
