#!/usr/bin/env python
"""Static SASS statistics per kernel of a built .so / .cubin:  python tools/sass_stats.py <file> [name-substring]
Prints the instruction count and opcode histogram for each matching kernel (cuobjdump -sass)."""
import collections
import re
import subprocess
import sys

path = sys.argv[1]
needle = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
cur, stats = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and cur:
        ins = m.group(1).strip()
        parts = ins.split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        stats[cur][op.split(".")[0]] += 1
for k, c in stats.items():
    if needle in k:
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        print(f"{name[:100]}: {sum(c.values())} instructions")
        print("   " + " ".join(f"{o}:{n}" for o, n in c.most_common(18)))
