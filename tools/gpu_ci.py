"""Run the GPU test-suite in isolated subprocesses (a trapped kernel poisons its CUDA context, so every group gets
its own process and timeout) and write logs + a summary under gpurun_out/ci/. Usage on the GPU box:

    python tools/gpu_ci.py [-k substring] [--each]      # --each: one process per test id
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-k", default="")
    ap.add_argument("--each", action="store_true")
    ap.add_argument("--split", default="test_linear_fwd_bwd", help="functions whose parametrisations run one per process")
    ap.add_argument("--timeout", type=int, default=300)
    ap.add_argument("paths", nargs="*", default=["tests"])
    a = ap.parse_args()
    out_dir = os.path.join(ROOT, "gpurun_out", "ci")
    os.makedirs(out_dir, exist_ok=True)
    col = subprocess.run([sys.executable, "-m", "pytest", *a.paths, "-m", "gpu", "--collect-only", "-q"], cwd=ROOT,
                         capture_output=True, text=True)
    ids = [l.strip() for l in col.stdout.splitlines() if "::" in l and (a.k in l)]
    groups = {}
    for i in ids:
        fn = i.split("[")[0]
        key = i if (a.each or fn.split("::")[-1] in a.split.split(",")) else fn
        groups.setdefault(key, []).append(i)
    summary = []
    for key, members in groups.items():
        name = re.sub(r"[^A-Za-z0-9_.-]+", "_", key)[-120:]
        log = os.path.join(out_dir, name + ".log")
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", *members, "-q", "-x", "--no-header", "-p", "no:cacheprovider"],
                               cwd=ROOT, capture_output=True, text=True, timeout=a.timeout)
            rc, text = r.returncode, r.stdout + "\n" + r.stderr
        except subprocess.TimeoutExpired as e:
            rc, text = -9, f"TIMEOUT after {a.timeout}s\n{e.stdout or ''}\n{e.stderr or ''}"
        with open(log, "w") as f:
            f.write(text)
        tail = [l for l in text.splitlines() if l.strip()][-1:] or [""]
        summary.append({"group": key, "n": len(members), "rc": rc, "sec": round(time.time() - t0, 1), "tail": tail[0][:200]})
        print(f"[{'ok' if rc == 0 else 'FAIL'}] {key} ({len(members)} tests, {summary[-1]['sec']}s) {tail[0][:160]}", flush=True)
        if rc != 0:
            errs = [l for l in text.splitlines() if re.search(r"Error|error|assert|rel err|mbarrier|TIMEOUT", l)][:12]
            print("    " + "\n    ".join(errs), flush=True)
    with open(os.path.join(out_dir, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    bad = [s for s in summary if s["rc"] != 0]
    print(f"{len(summary) - len(bad)}/{len(summary)} groups passed")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
