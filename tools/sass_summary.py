"""Per-kernel SASS opcode evidence of libmtlora_b200.so (cuobjdump -sass): which kernels carry tcgen05 (UTCHMMA / UTCBAR),
TMEM loads (LDTM), TMA (UTMALDG / UTMASTG / UBLKCP), mma.sync (HMMA), ldmatrix (LDSM), cp.async (LDGSTS).
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mtlora_b200", "libmtlora_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "HMMA", "LDSM", "LDGSTS", "SYNCS",
         "MUFU", "FFMA2", "RED", "ATOMG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = kern.replace("mtl::(anonymous namespace)::", "").replace("void ", "")
            kern = re.sub(r"\((?!anonymous).*", "", kern)
            counts[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and kern:
            op = m.group(1)
            total[kern] += 1
            for w in WATCH:
                if op.startswith(w):
                    counts[kern][w] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS opcode counts per kernel (static instruction counts, sm_100a)")
    print(f"# {'kernel':70s} {'instrs':>7s}  " + " ".join(f"{w:>7s}" for w in WATCH))
    for k, c in counts.items():
        if total[k] == 0:
            continue
        print(f"{k[:72]:72s} {total[k]:7d}  " + " ".join(f"{c.get(w, 0):7d}" for w in WATCH))


if __name__ == "__main__":
    main()
