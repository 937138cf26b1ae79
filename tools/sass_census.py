"""Per-kernel SASS census of libumhs_sm100.so: the mnemonics that prove (or disprove) a Blackwell-native kernel
(B200_PROFILING.md): UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG /
UBLKCP = TMA, HMMA = legacy mma.sync, REDG / ATOMG = global reductions, LDGMC / multimem = NVLS, LDGSTS = cp.async.
    python tools/sass_census.py [lib.so] > profiles/r2_sass_census.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "unsupervised-hyperspectral-nerf_b200/umhs_b200/libumhs_sm100.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "REDG", "ATOMG", "LDGMC", "STGMC",
        "LDGSTS", "BAR.SYNC", "BAR.ARV", "SYNCS", "MUFU", "SHFL", "LDG", "STG", "LDS", "STS", "FFMA"]
per = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        per[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        per[name]["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                per[name][k] += 1
dem = subprocess.run(["c++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
print("# SASS census per kernel (`cuobjdump -sass`, sm_100a)\n")
print("`UTCHMMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st, `UTCBAR` = tcgen05.commit, `UTMALDG`/`UTMASTG`/`UBLKCP` = "
      "TMA, `HMMA` = legacy mma.sync (must be 0), `REDG` = red.global, `LDGMC`/`STGMC` = multimem (NVLS).\n")
print("| kernel | instr | " + " | ".join(KEYS) + " |")
print("|---|---:|" + "---:|" * len(KEYS))
tot = collections.Counter()
for (mangled, c), d in zip(per.items(), dem):
    short = re.sub(r"\(.*", "", d).replace("void ", "").replace("umhs::", "")[:58]
    print(f"| `{short}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
    tot.update(c)
print(f"| **all kernels** | {tot['_total']} | " + " | ".join(str(tot[k]) for k in KEYS) + " |")
