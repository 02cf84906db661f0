"""Summarise an .ncu-rep (read offline with `ncu -i`): headline metrics, SASS opcode mix, top stall sites.
    python tools/ncu_summary.py gpurun_out/prof_fc1_fwd.ncu-rep [--json out.json]"""
import collections
import csv
import io
import json
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__inst_executed.avg.per_cycle_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__cycles_active.avg', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_bytes.sum', 'launch__occupancy_limit_registers',
        'launch__shared_mem_per_block_dynamic']


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = {}
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    h, u, v = rows[0], rows[1], rows[2]
    out["kernel"] = v[h.index("Kernel Name")] if "Kernel Name" in h else ""
    for i, n in enumerate(h):
        if n in WANT:
            out[n] = f"{v[i]} {u[i]}".strip()
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    hdr, data = rows[hi], rows[hi + 1:]
    iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    byop, stalls = collections.Counter(), collections.Counter()
    tot = 0
    for r in data:
        try:
            e = int(r[iE])
        except ValueError:
            continue
        toks = r[iS].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        byop[op] += e
        tot += e
        for i in stall_cols:
            stalls[hdr[i]] += int(r[i] or 0)
    out["sass_inst_total"] = tot
    out["sass_mix_pct"] = {k: round(100 * c / tot, 1) for k, c in byop.most_common(14)}
    ssum = sum(stalls.values())
    out["stall_pct"] = {k: round(100 * c / ssum, 1) for k, c in stalls.most_common(8)}
    top = sorted(data, key=lambda r: -int(r[iSm] or 0))[:8]
    out["top_stall_sites"] = [{"samples": int(r[iSm]), "executed": int(r[iE]), "sass": r[iS].strip()[:80]} for r in top]
    print(json.dumps(out, indent=1))
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
