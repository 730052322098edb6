#!/usr/bin/env python
"""Split an `ncu --page source --csv` dump of ONE kernel into the regions between BAR.SYNC instructions (SASS order) and
print, per region, the warp instructions executed per unit of work and the share of the stall samples with its top
reasons.  usage: python scripts/ncu_regions.py gpurun_out/TAG_ncu/source.csv [units]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter()
ops = collections.Counter()
region = 0
regs = collections.defaultdict(collections.Counter)
rinstr = collections.Counter()
rops = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix['Source']]
    if 'BAR.SYNC' in src:
        region += 1
    n = int(r[ix['Instructions Executed']] or 0)
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = (m.group(2) if m else src[:10]).split('.')[0]
    ops[op] += n
    rops[region][op] += n
    rinstr[region] += n
    for c in cols:
        v = int(r[ix[c]] or 0)
        tot[c] += v
        regs[region][c] += v
S = sum(tot.values()) or 1
T = sum(ops.values())
print(f"warp instructions per unit: {T / units:.0f}; samples {S}")
print("ops: " + ", ".join(f"{k} {v / units:.0f}" for k, v in ops.most_common(14)))
print("stalls: " + ", ".join(f"{c[6:]} {100 * v / S:.1f}%" for c, v in tot.most_common(8)))
for k in sorted(regs):
    s = sum(regs[k].values())
    top = ", ".join(f"{c[6:]} {100 * v / s:.0f}%" for c, v in regs[k].most_common(4)) if s else ""
    topo = ", ".join(f"{o} {v / units:.0f}" for o, v in rops[k].most_common(5))
    print(f"  R{k:2d} instr {rinstr[k] / units:8.0f}  samples {100 * s / S:5.1f}%  [{top}]  {topo}")
