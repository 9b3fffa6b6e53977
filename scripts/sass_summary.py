#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-native SASS opcodes in libao_b200.so (tcgen05.mma = UTC*MMA, tcgen05.ld / st =
LDTM / STTM, tcgen05.cp = UTCCP, tcgen05.commit = UTCBAR, TMA = UTMALDG / UBLKCP; HMMA would be the legacy mma.sync
path) -> profiles/r02_sass_summary.txt.   python scripts/sass_summary.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ("UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTCCP", "UTCBAR", "UTMALDG", "UTMAPF", "UBLKCP", "UBLKPF", "HMMA", "LDGSTS")
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "ao_b200", "lib", "libao_b200.so")], capture_output=True, text=True).stdout
counts, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    m = re.search(r"\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]+)", line)
    if m and name and m.group(1) in OPS:
        counts[name][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
with open(os.path.join(ROOT, "profiles", "r02_sass_summary.txt"), "w") as f:
    f.write("# cuobjdump -sass ao_b200/lib/libao_b200.so: tcgen05 / TMEM / TMA opcodes per kernel (scripts/sass_summary.py)\n")
    for (mangled, c), nice in zip(counts.items(), names):
        nice = nice.replace("CUtensorMap_st", "TMap")
        f.write(f"{nice[:150]}\n    " + ("  ".join(f"{k}={v}" for k, v in sorted(c.items())) or "(CUDA-core kernel)") + "\n")
print(open(os.path.join(ROOT, "profiles", "r02_sass_summary.txt")).read()[:3000])
