"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the built library (cuobjdump -sass):
UTC*MMA = tcgen05.mma, UTMALDG = TMA tensor loads, UBLKCP = 1-D bulk copies, LDTM/STTM = tcgen05.ld/st, HMMA = mma.sync.
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "semstereo_b200", "libsemstereo_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "FFMA", "MUFU")
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  ({len(counts)} kernels); instruction counts per kernel")
print("# " + " ".join(f"{k:>8}" for k in KEYS) + "  kernel")
tot = collections.Counter()
for (k, c), n in zip(counts.items(), names):
    tot.update(c)
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    n = re.sub(r"\(.*", "", n)
    print("  " + " ".join(f"{c.get(x, 0):>8}" for x in KEYS) + "  " + n[:110])
print("# " + " ".join(f"{tot.get(x, 0):>8}" for x in KEYS) + "  TOTAL")
