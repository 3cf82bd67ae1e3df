"""SASS opcode census of libsct_b200.so: per kernel, how many tcgen05 / TMEM / TMA instructions it contains
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store /
reduce, UBLKCP = 1-D bulk copy, SYNCS = mbarrier) and that no legacy tensor path (HMMA / HGMMA) is present.
    python tools/sass_census.py [path/to/libsct_b200.so] > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sct_gan_b200", "libsct_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "FFMA2", "FMUL2",
       "FADD2", "MUFU.EX2", "MUFU.TANH", "HMMA", "HGMMA"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"sct::\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        cur = per.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    cur["_total"] += 1
    for o in OPS:
        if o == "UTCHMMA.2CTA":
            if op.startswith("UTCHMMA") and ".2CTA" in line:
                cur[o] += 1
        elif op == o or op.startswith(o + "."):
            cur[o] += 1
print(f"# SASS census of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a)")
print(f"{'kernel':72s} {'instr':>7s} " + " ".join(f"{o:>8s}" for o in OPS))
tot = collections.Counter()
for name, c in per.items():
    if not any(c[o] for o in OPS):
        continue
    print(f"{name[:72]:72s} {c['_total']:7d} " + " ".join(f"{c[o]:8d}" for o in OPS))
    tot.update(c)
print(f"{'TOTAL (kernels listed)':72s} {tot['_total']:7d} " + " ".join(f"{tot[o]:8d}" for o in OPS))
assert tot["HMMA"] == 0 and tot["HGMMA"] == 0, "legacy tensor-core path found"
