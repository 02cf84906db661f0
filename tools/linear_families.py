"""Per-family table of the fused-linear launches of one step from a bench --profile-ops file: time, TFLOP/s (dense + true-rank
adapter flops) and algorithmic GB/s (SURVEY.md §8d formula) per (pass, M, K, N, streams, rank space).
    python tools/linear_families.py profiles/r02_ops_profile.json > profiles/r02_linear_families.txt"""
import collections
import json
import sys


def main(path):
    d = json.load(open(path))
    agg = collections.OrderedDict()
    for c in d["calls"]:
        if c["name"] in ("mtl_linear_fwd", "mtl_linear_bwd_input", "mtl_linear_bwd_params") and c["meta"]:
            k = (c["name"][11:],) + tuple(c["meta"][1:8])
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += c["ms"]
    n_steps = d.get("steps_profiled", 2)
    print(f"# {d['workload']}")
    print("# CUDA-event time per C-ABI call (launches of < 20 us are dominated by the event overhead of the profiled run)")
    print(f"# {'pass':10s} {'M':>7s} {'K':>5s} {'N':>5s} s_in s_out R_pad r | calls/step ms/step  us/call  TFLOP/s  alg GB/s  frac of 6551")
    tot = 0.0
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        kind, M, K, N, sa, sb, R, rs = k
        fl = 2.0 * M * K * N + 2.0 * M * rs * (K + N)
        if kind == "fwd":
            by = 2 * (M * K * sa + N * K + rs * (K + N) + M * N * sb)
        elif kind == "bwd_input":
            by = 2 * (M * N * sb + N * K + rs * (K + N) + M * K * sa)
        else:
            by = 2 * (M * K * sa + M * N * sb + 2 * M * R) + 4 * rs * (K + N)
            fl = 2.0 * M * R * (K + N)
        per = ms / n
        gbs = by / per / 1e6
        print(f"{kind:12s} {M:7d} {K:5d} {N:5d} {sa:4d} {sb:5d} {R:5d} {rs:3d} | {n / n_steps:8.0f} {ms / n_steps:8.3f} {1e3 * per:8.1f} "
              f"{fl / per / 1e9:8.1f} {gbs:9.0f} {gbs / 6551:8.3f}")
        tot += ms / n_steps
    print(f"# total {tot:.3f} ms/step")


if __name__ == "__main__":
    main(sys.argv[1])
