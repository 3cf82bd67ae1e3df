"""Turn the scratch output of tools/capture_profiles.sh (gpurun_out/r02/) into the tracked files under profiles/.

    python tools/summarise_profiles.py [TAG]      (TAG defaults to r02; needs ncu on PATH for the .ncu-rep exports)

Writes  profiles/TAG_bench_*.json, TAG_launches.csv(.gz), TAG_launch_shares.txt, TAG_gemm_traffic.json,
        TAG_ncu_{gemm,attn}_full_raw.csv, TAG_{attn,gemm,ffn,rowwise}_bench.txt, TAG_trace_summary.txt."""
import collections
import csv
import gzip
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
SRC = os.path.join(ROOT, "gpurun_out", TAG)
DST = os.path.join(ROOT, "profiles")


def family(name):
    m = re.search(r"sct::(?:<unnamed>::)?(\w+)", name)
    if not m:
        return "torch: " + re.sub(r"^void ", "", name)[:60]
    k = m.group(1)
    if k == "gemm_kernel":
        t = re.search(r"gemm_kernel<([^>]*)>", name)
        return "gemm_kernel<" + (t.group(1).replace(" ", "") if t else "") + ">"
    return k


def launches():
    path = os.path.join(SRC, "launches.csv")
    if not os.path.exists(path):
        return
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        elif unit in ("us", "usecond"):
            v *= 1e3
        elif unit in ("ms", "msecond"):
            v *= 1e6
        d[r["Metric Name"]] = v
    fam = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in per.values():
        f = fam[family(d["name"])]
        f[0] += 1
        f[1] += d.get("gpu__time_duration.sum", 0.0)
        f[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot = sum(f[1] for f in fam.values())
    with open(os.path.join(DST, f"{TAG}_launch_shares.txt"), "w") as out:
        out.write(f"# one train step under ncu (cold cache, serialised): {len(per)} launches, {tot / 1e6:.2f} ms of kernels\n")
        out.write("# command: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                  "--clock-control none --profile-from-start off --csv python bench.py --ncu-step --no-cpu-baseline\n")
        out.write(f"{'family':64s} {'launches':>8s} {'ms':>9s} {'share':>7s} {'DRAM MB/launch':>15s}\n")
        for k, f in sorted(fam.items(), key=lambda kv: -kv[1][1]):
            if f[1] / tot < 0.001:
                continue
            out.write(f"{k:64s} {f[0]:8d} {f[1] / 1e6:9.3f} {f[1] / tot:7.1%} {f[2] / f[0] / 1e6:15.2f}\n")
    with gzip.open(os.path.join(DST, f"{TAG}_launches.csv.gz"), "wt") as out:
        out.write("".join(lines))
    g = [f for k, f in fam.items() if k.startswith("gemm_kernel")]
    n = sum(f[0] for f in g)
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    rec = {"kernel": "gemm_kernel (all variants)", "launches": n, "ms_total_under_ncu": sum(f[1] for f in g) / 1e6,
           "dram_bytes_total": sum(f[2] for f in g), "dram_bytes_per_launch": sum(f[2] for f in g) / max(n, 1),
           "commit": commit,
           "command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                      "--profile-from-start off --csv python bench.py --ncu-step --no-cpu-baseline",
           "source": f"profiles/{TAG}_launches.csv.gz"}
    with open(os.path.join(DST, f"{TAG}_gemm_traffic.json"), "w") as out:
        json.dump(rec, out, indent=1)
    print("launch list:", len(per), "launches,", f"{tot / 1e6:.2f} ms; GEMM DRAM/launch {rec['dram_bytes_per_launch'] / 1e6:.1f} MB")


def ncu_raw(rep, dst):
    path = os.path.join(SRC, rep)
    if not os.path.exists(path):
        return
    r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    if r.returncode == 0:
        with open(os.path.join(DST, dst), "w") as out:
            out.write(r.stdout)
        print("wrote", dst)
    else:
        print("ncu export failed:", r.stderr[-300:])


def main():
    for f in sorted(os.listdir(SRC)):
        p = os.path.join(SRC, f)
        if f.startswith("bench_") and f.endswith(".json") and os.path.getsize(p) > 0:
            shutil.copy(p, os.path.join(DST, f"{TAG}_{f}"))
        if f.endswith("_bench.txt") and os.path.getsize(p) > 0:
            shutil.copy(p, os.path.join(DST, f"{TAG}_{f}"))
    if os.path.exists(os.path.join(SRC, "gpu_tests.log")):
        with open(os.path.join(SRC, "gpu_tests.log")) as f:
            tail = f.read().strip().splitlines()[-3:]
        with open(os.path.join(DST, f"{TAG}_gpu_tests.txt"), "w") as out:
            out.write("python -m pytest tests -m gpu -q   (one B200)\n" + "\n".join(tail) + "\n")
    launches()
    ncu_raw("gemm_prof.ncu-rep", f"{TAG}_ncu_gemm_full_raw.csv")
    ncu_raw("attn_prof.ncu-rep", f"{TAG}_ncu_attn_full_raw.csv")
    tr = os.path.join(SRC, "trace.json.gz")
    if os.path.exists(tr):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "trace_summary.py"), tr], capture_output=True, text=True)
        with open(os.path.join(DST, f"{TAG}_trace_summary.txt"), "w") as out:
            out.write(r.stdout)


if __name__ == "__main__":
    main()
