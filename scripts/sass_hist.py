#!/usr/bin/env python
"""Instruction histogram of one kernel from `cuobjdump -sass` output (static counts).
usage: cuobjdump -sass lib.so | python scripts/sass_hist.py <kernel-name-substring>"""
import collections
import re
import sys

name = sys.argv[1]
cur, hist, total = None, collections.Counter(), 0
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None or name not in cur:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        hist[m.group(1)] += 1
        total += 1
print(total, "instructions")
for k, v in hist.most_common(25):
    print(f"{v:6d} {k}")
