#!/usr/bin/env python3
"""Attribute the warp-instructions of one kernel in an .ncu-rep to CUDA source lines.

ncu's CSV source page carries per-SASS-instruction counters but no line numbers; nvdisasm -g on the
cubin of the same build carries the line table.  The two listings have the same instruction order,
so they are zipped.   usage: ncu_lines.py <rep> <kernel-substring> <object.o> [top]
"""
import collections
import csv
import re
import subprocess
import sys

rep, kname, obj = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
cur, kernels = None, []
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        kernels.append(cur)
    elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and r:
        cur['rows'].append(r)
k = [k for k in kernels if kname in k['name']][0]
h = k['hdr']
ie, src = h.index('Instructions Executed'), h.index('Source')
smp = h.index('# Samples')
import os, tempfile
if not obj.endswith('.cubin'):
    d = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=d, capture_output=True)
    obj = os.path.join(d, [f for f in os.listdir(d) if f.endswith('.cubin')][0])
dis = subprocess.run(['nvdisasm', '-g', '-c', obj], capture_output=True, text=True).stdout
# walk the disassembly of the matching function
lines, cur_line, inside = [], None, False
for l in dis.splitlines():
    if l.startswith('.text.') or re.match(r'\s*\.section\s+\.text\.', l):
        inside = kname in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur_line)
n = min(len(lines), len(k['rows']))
if len(lines) != len(k['rows']):
    print('warning: %d disassembled vs %d profiled instructions' % (len(lines), len(k['rows'])))
agg, samp = collections.Counter(), collections.Counter()
tot = 0
for i in range(n):
    c = int(k['rows'][i][ie] or 0)
    agg[lines[i]] += c
    samp[lines[i]] += int(k['rows'][i][smp] or 0)
    tot += c
stot = sum(samp.values()) or 1
print(k['name'][:70], 'warp-instructions', tot)
cache = {}
for (key, c) in agg.most_common(top):
    txt = ''
    if key:
        f, ln = key
        try:
            if f not in cache:
                import glob
                cand = glob.glob('/root/repo/orb_slam_fusion_b200/csrc/' + f)
                cache[f] = open(cand[0]).read().splitlines() if cand else []
            txt = cache[f][ln - 1].strip()[:90] if cache[f] else ''
        except Exception:
            pass
    print("%6.2f%% inst %6.2f%% stall-samples  %-22s %s" % (100.0 * c / tot, 100.0 * samp[key] / stot, '%s:%d' % key if key else '?', txt))
