"""Copy the summaries of gpurun_out/evidence (tools/evidence.sh) into profiles/ under a round prefix.
    python tools/evidence_collect.py r02"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EV = os.path.join(ROOT, "gpurun_out", "evidence")
PR = os.path.join(ROOT, "profiles")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    os.makedirs(PR, exist_ok=True)
    for src, dst in (("bench.json", f"{tag}_bench.json"), ("bench_reference.json", f"{tag}_bench_reference.json"),
                     ("bench_reference_gpu.json", f"{tag}_bench_reference_gpu.json"), ("ops_profile.json", f"{tag}_ops_profile.json"),
                     ("attn_ab.txt", f"{tag}_attn_ab.txt"), ("linear_ab.txt", f"{tag}_linear_ab.txt")):
        p = os.path.join(EV, src)
        if os.path.exists(p) and os.path.getsize(p) > 0:
            shutil.copy(p, os.path.join(PR, dst))
    # test summary
    log = os.path.join(EV, "gpu_tests.log")
    if os.path.exists(log):
        tail = [l for l in open(log).read().splitlines() if l.strip()][-1]
        smoke = open(os.path.join(EV, "smoke.log")).read().splitlines()[-1] if os.path.exists(os.path.join(EV, "smoke.log")) else ""
        json.dump({"pytest -m gpu": tail, "smoke": smoke}, open(os.path.join(PR, f"{tag}_gpu_tests_summary.json"), "w"), indent=1)
    # launch list -> per-kernel summary + DRAM traffic of the linear kernel
    csv = os.path.join(EV, "launches.csv")
    if os.path.exists(csv):
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launches_summary.py"), csv,
                        os.path.join(PR, f"{tag}_launches.json")], check=False)
        d = json.load(open(os.path.join(PR, f"{tag}_launches.json")))
        lin = [k for k in d["kernels"] if k["kernel"].startswith("mtl_linear_kernel")]
        n = sum(k["launches"] for k in lin)
        rd = sum(k.get("dram__bytes_read.sum", 0.0) for k in lin)
        wr = sum(k.get("dram__bytes_write.sum", 0.0) for k in lin)
        if n:
            json.dump({"dram_bytes_per_launch": (rd + wr) / n, "launches": n, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "source": f"profiles/{tag}_launches.json: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the "
                                 "mtl_linear_kernel launches of ONE bench step (bench.py --ncu-range); average per launch, like "
                                 "roofline.achieved"}, open(os.path.join(PR, "linear_traffic.json"), "w"), indent=1)
    # ncu --set full summaries
    for f in sorted(os.listdir(EV)):
        if f.endswith(".ncu-rep"):
            name = f[:-len(".ncu-rep")].replace("ncu_", "")
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(EV, f), "--json",
                            os.path.join(PR, f"{tag}_ncu_{name}.json")], capture_output=True)
    print(sorted(x for x in os.listdir(PR) if x.startswith(tag)))


if __name__ == "__main__":
    main()
