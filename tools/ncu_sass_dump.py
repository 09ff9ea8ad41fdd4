#!/usr/bin/env python3
"""Print the SASS of one kernel with executed warp-instruction counts and source lines.
usage: ncu_sass_dump.py sass.csv lib.sass kernel_substr[|mangled_substr] [file:lo-hi]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1]))); want, _, want_sass = sys.argv[3].partition('|')
want_sass = want_sass or want
ks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'hdr': None, 'rows': []}; ks.append(cur)
    elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address': cur['hdr'] = r
    elif cur is not None and cur['hdr'] and r: cur['rows'].append(r)
k = [x for x in ks if want in x['name']][0]
h = k['hdr']; ie = h.index('Instructions Executed'); isrc = h.index('Source'); ismp = h.index('# Samples')
lines, cur, infn = [], None, False
for l in open(sys.argv[2]):
    if l.startswith('.text.') or re.match(r'\s*\.section\s+\.text\.', l): infn = want_sass in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m and infn: lines.append((int(m.group(1), 16), cur, m.group(2)))
flt = None
if len(sys.argv) > 4:
    f, lh = sys.argv[4].split(':'); lo, hi = lh.split('-'); flt = (f, int(lo), int(hi))
for r, (off, src, txt) in zip(k['rows'], lines):
    if flt and not (src and src[0] == flt[0] and flt[1] <= src[1] <= flt[2]): continue
    print("%05x %9s %5s  %-22s %s" % (off, r[ie], r[ismp], '%s:%d' % src if src else '', txt))
