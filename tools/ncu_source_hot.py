#!/usr/bin/env python3
"""Per-source-line instruction counts of the first kernel in an .ncu-rep (needs -lineinfo):
correlates the SASS view (instructions executed, stall samples) back to CUDA source lines."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda' if False else 'cuda,sass'],
                     capture_output=True, text=True).stdout
# fall back: the combined view is not available on every ncu; use two passes
sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
kernels = []
cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        kernels.append(cur)
    elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and r:
        cur['rows'].append(r)
k = kernels[0]
h = k['hdr']
ie, ts, src = h.index('Instructions Executed'), h.index('# Samples'), h.index('Source')
tot = sum(int(r[ie] or 0) for r in k['rows'])
print(k['name'][:80], 'SASS instructions', len(k['rows']), 'warp-instructions executed', tot)
top = sorted(k['rows'], key=lambda r: -int(r[ie] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for r in top:
    print("%8.3f%%  samples %6s  %s" % (100.0 * int(r[ie] or 0) / tot, r[ts], r[src][:100]))
