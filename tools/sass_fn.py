#!/usr/bin/env python
"""Print the SASS of one kernel of a built .so:  python tools/sass_fn.py <file> <mangled-name-substring> [out]"""
import re, subprocess, sys
path, needle = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
cur, keep, lines = None, False, []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        keep = needle in m.group(1)
        if keep:
            lines.append(line)
        continue
    if keep and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        lines.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line))
text = "\n".join(lines)
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(text)
else:
    print(text)
