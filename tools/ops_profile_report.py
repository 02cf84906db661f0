"""Summarise a bench.py --profile-ops JSON: per (entry point, shape) time per step and achieved algorithmic GB/s."""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import linear_alg_bytes  # noqa: E402


def main(path, top=40):
    d = json.load(open(path))
    steps = d.get("steps_profiled", 2)
    agg = collections.OrderedDict()
    for c in d["calls"]:
        k = (c["name"], tuple(c["meta"]) if c["meta"] else None)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += c["ms"]
    rows = []
    for (n, m), (cnt, ms) in agg.items():
        gbs = linear_alg_bytes(m) * cnt / ms / 1e6 if m else None
        rows.append((ms / steps, n, m, cnt // steps, gbs))
    rows.sort(key=lambda r: -r[0])
    print(f"total {sum(r[0] for r in rows):.2f} ms/step over {sum(r[3] for r in rows)} calls/step")
    for r in rows[:top]:
        print(f"{r[0]:8.3f} ms/step  x{r[3]:3d} {r[1]:26s} {r[2]}  {'%.0f GB/s' % r[4] if r[4] else ''}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
