"""Count the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st ->
LDTM / STTM, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS, red.global -> REDG; legacy mma.sync would show HMMA) per kernel of
libtensoflow_b200.so.  Runs without a GPU:  python scripts/sass_evidence.py > profiles/r1_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "tensoflow_b200", "libtensoflow_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UBLKCP|UTMALDG|UTMASTG|SYNCS|REDG|RED|HMMA|HGMMA)\b")
counts = collections.OrderedDict()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(anonymous namespace\)::", "", fn).split("(")[0]
        counts.setdefault(fn, collections.Counter())
        continue
    if fn:
        for op in pat.findall(line.split("/*")[1] if "/*" in line and line.strip().startswith("/*") else line):
            counts[fn][op] += 1
print("# cuobjdump -sass tensoflow_b200/libtensoflow_b200.so (sm_100a), mnemonic counts per kernel; kernels without any are omitted")
for fn, c in sorted(counts.items()):
    if c:
        print(f"{fn:48s} " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
if any("HMMA" == k or "HGMMA" == k for c in counts.values() for k in c):
    print("# WARNING: legacy tensor-core mnemonics present", file=sys.stderr)
