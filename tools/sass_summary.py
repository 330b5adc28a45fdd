#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that identify the Blackwell-native paths (B200_PROFILING.md):
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk,
UTCBAR = tcgen05.commit, HMMA = mma.sync, LDGSTS = cp.async.   usage: tools/sass_summary.py [lib.so] > profiles/…md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "geomae_b200/libgeomae_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "HMMA", "LDGSTS", "REDG", "RED"]
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
                break
print(f"SASS mnemonic counts per kernel of `{lib}` (cuobjdump -sass, sm_100a)\n")
print("| kernel | " + " | ".join(KEYS) + " |")
print("|---|" + "---:|" * len(KEYS))
for name, c in counts.items():
    if any(c[k] for k in KEYS[:8]):
        print(f"| `{name}` | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
