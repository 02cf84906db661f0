"""Aggregate an `ncu --metrics ... --csv` launch list by kernel: launches, total metric value, share of the time.
    python tools/launches_summary.py gpurun_out/launches.csv profiles/r01_launches.json"""
import collections
import csv
import json
import re
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("mtl::<unnamed>::", "").replace("void ", "")
    return name[:90]


def main(path, out):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    per = collections.OrderedDict()
    for r in rd:
        k = short(r["Kernel Name"])
        m = r["Metric Name"]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if m == "gpu__time_duration.sum":
            v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)   # -> us
        d = per.setdefault(k, collections.defaultdict(float))
        d[m] += v
        if m == "gpu__time_duration.sum":
            d["launches"] += 1
    tot = sum(d["gpu__time_duration.sum"] for d in per.values()) or 1.0
    res = []
    for k, d in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        e = {"kernel": k, "launches": int(d["launches"]), "time_us": round(d["gpu__time_duration.sum"], 1),
             "share": round(d["gpu__time_duration.sum"] / tot, 4)}
        for m, v in d.items():
            if m.startswith("dram__"):
                e[m] = v
        res.append(e)
    summary = {"source": path, "note": "ncu launch list (cold-cache, serialised): compare SHARES, not absolutes",
               "total_time_us": round(tot, 1), "kernels": res}
    with open(out, "w") as f:
        json.dump(summary, f, indent=1)
    for e in res[:14]:
        print(f"{e['share']*100:5.1f}%  x{e['launches']:4d}  {e['time_us']:10.1f} us  {e['kernel']}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
