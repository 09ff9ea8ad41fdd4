#!/usr/bin/env python3
"""Join an ncu source-page CSV (per-SASS-instruction executed counts) with nvdisasm --print-line-info of the
linked library and print executed warp-instructions per source line / per line range.
usage: ncu_phase.py sass.csv lib.sass kernel_substr[|mangled_substr] [file.cu:lo-hi=name ...]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
want, _, want_sass = sys.argv[3].partition('|')
want_sass = want_sass or want
ks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'hdr': None, 'rows': []}; ks.append(cur)
    elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address': cur['hdr'] = r
    elif cur is not None and cur['hdr'] and r: cur['rows'].append(r)
k = [x for x in ks if want in x['name']][0]
h = k['hdr']; ie = h.index('Instructions Executed'); isrc = h.index('Source')
# nvdisasm listing: find the function
lines, cur, infn = [], None, False
for l in open(sys.argv[2]):
    if l.startswith('.text.') or re.match(r'\s*\.section\s+\.text\.', l): infn = want_sass in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m and infn: lines.append((int(m.group(1), 16), cur, m.group(2)))
assert len(lines) == len(k['rows']), (len(lines), len(k['rows']))
per = collections.Counter(); tot = 0
for r, (off, src, txt) in zip(k['rows'], lines):
    n = int(r[ie] or 0); tot += n; per[src] += n
ranges = []
for a in sys.argv[4:]:
    spec, name = a.split('='); f, lh = spec.split(':'); lo, hi = lh.split('-'); ranges.append((f, int(lo), int(hi), name))
print(k['name'][:60], 'warp-instructions', tot)
if ranges:
    ph = collections.Counter()
    for (f, l), n in per.items():
        for rf, lo, hi, name in ranges:
            if f == rf and lo <= l <= hi: ph[name] += n; break
        else: ph['%s:%d' % (f, l)] += n
    for p, n in sorted(ph.items(), key=lambda x: -x[1])[:40]: print("%6.2f%%  %s" % (100 * n / tot, p))
else:
    for (f, l), n in sorted(per.items(), key=lambda x: -x[1])[:60]: print("%6.2f%%  %s:%d" % (100 * n / tot, f, l))
