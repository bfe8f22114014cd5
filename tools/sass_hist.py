#!/usr/bin/env python3
"""tools/sass_hist.py <lib.so> <mangled-kernel-prefix> -- opcode histogram of one kernel's SASS."""
import collections
import re
import subprocess
import sys

lib, name = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
secs = re.split(r"\n\s*Function : ", txt)
sec = [x for x in secs if x.startswith(name)][0]
ops = collections.Counter()
n = 0
for line in sec.splitlines():
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        ops[m.group(2).split(".")[0]] += 1
        n += 1
print("total", n)
for k, v in ops.most_common(40):
    print(k, v)
