"""Instruction counts per kernel from `cuobjdump -sass` of the built library (no GPU needed):
which kernels use the TMA engine (UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk), mbarriers
(SYNCS), cp.async (LDGSTS), packed fp32x2 FMAs (FFMA2) ...  usage: sass_evidence.py [lib] [out]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "somax_b200/lib/libsomax_b200.so"
out_path = sys.argv[2] if len(sys.argv) > 2 else "profiles/r01_sass_evidence.txt"
keys = ["UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "FFMA2", "DFMA", "F2F", "SHFL", "LDS", "STS", "LDG", "STG", "MEMBAR", "FENCE"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts, cur = collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None or "/*" not in line:
        continue
    for k in keys:
        if re.search(r"\b" + k + r"\b", line):
            counts[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
rows = sorted((re.sub(r"\(.*", "", n), counts[f]) for n, f in zip(names, counts))
lines = ["# SASS instruction counts per kernel (cuobjdump -sass %s, sm_100a; tools/sass_evidence.py)" % lib,
         "# UTMALDG = cp.async.bulk.tensor (TMA tensor-map load), UBLKCP = cp.async.bulk (TMA engine, 1-D), SYNCS = mbarrier,",
         "# LDGSTS = cp.async, FFMA2 = packed fp32x2 FMA.  Only kernels that use one of those, or belong to the slab model.",
         "kernel | " + " | ".join(keys)]
for short, c in rows:
    if any(c[k] for k in ("UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "FFMA2")) or "seg_copy" in short or "slab_barrier" in short:
        lines.append(short[:120] + " | " + " | ".join(str(c[k]) for k in keys))
open(out_path, "w").write("\n".join(lines) + "\n")
print(len(rows), "kernels,", len(lines) - 4, "listed ->", out_path)
