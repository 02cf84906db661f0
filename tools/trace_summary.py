"""Summarise MTL_LINEAR_TRACE timelines: per-role spans and epilogue item durations (wait->got->done)."""
import collections
import sys

for f in sys.argv[1:]:
    lines = open(f).read().split("\n")
    print("====", f, lines[0])
    ev = collections.defaultdict(list)
    for l in lines[1:]:
        if l.strip():
            r, t, c = l.split()
            ev[int(r)].append((int(t), int(c)))
    for r in sorted(ev):
        e = ev[r]
        print(f" role {r}: {len(e)} events, span {(e[-1][0] - e[0][0]) / 1e3:.1f} us")
    for r in (2, 3):
        wait = got = None
        waits, items = [], collections.defaultdict(list)
        for t, c in ev.get(r, []):
            k = c // 1000000
            if k == 5:
                wait = t
            elif k == 6:
                got = t
                waits.append((got - wait) / 1e3)
            elif k == 7 and got is not None:
                items[c % 10].append((t - got) / 1e3)
        # inside a half: 6/9.5 -> 9 = waiting for slabs / inputs, 9 -> 9.5 = arithmetic
        last = None
        w_slab, w_math = [], []
        for t, c in ev.get(r, []):
            k = c // 500000
            if k in (12, 19):       # 6xxxxxx got accumulator, 95xxxxx half computed
                last = t
            elif k == 18 and last is not None:   # 9000000 slabs ready
                w_slab.append((t - last) / 1e3)
                last = t
            if k == 19 and w_slab:
                pass
        t9 = None
        for t, c in ev.get(r, []):
            if c // 500000 == 18:
                t9 = t
            elif c // 500000 == 19 and t9 is not None:
                w_math.append((t - t9) / 1e3)
        if w_slab:
            print(f" group {r - 2}: per half: wait for slabs/inputs {sum(w_slab) / len(w_slab):.2f} us, arithmetic {sum(w_math) / max(len(w_math), 1):.2f} us")
        if waits:
            print(f" group {r - 2}: mean wait-for-accumulator {sum(waits) / len(waits):.2f} us; item time by stream:",
                  {j: round(sum(v) / len(v), 2) for j, v in sorted(items.items())})
    # MMA issuer: chunk cadence
    ch = [t for t, c in ev.get(1, []) if c // 1000000 == 4]
    if len(ch) > 2:
        d = [(b - a) / 1e3 for a, b in zip(ch, ch[1:])]
        d.sort()
        print(f" MMA chunk cadence: median {d[len(d) // 2]:.2f} us over {len(d)} chunks")
    w = [t for t, c in ev.get(1, []) if c // 1000000 == 1]
    if len(w) > 2:
        d = [(b - a) / 1e3 for a, b in zip(w, w[1:])]
        print(f" work-item period: mean {sum(d) / len(d):.2f} us over {len(d)} items")
