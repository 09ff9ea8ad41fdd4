#!/usr/bin/env python3
"""Executed warp-instructions by opcode for every kernel of an ncu source-page CSV
(ncu -i rep --page source --csv --print-source sass > sass.csv).  usage: ncu_opmix.py sass.csv [kernel_substr]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
ks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'hdr': None, 'rows': []}; ks.append(cur)
    elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address': cur['hdr'] = r
    elif cur is not None and cur['hdr'] and r: cur['rows'].append(r)
seen = set()
for k in ks:
    if want not in k['name'] or not k['rows']: continue
    h = k['hdr']; ie = h.index('Instructions Executed'); isrc = h.index('Source')
    c = collections.Counter(); tot = 0
    for r in k['rows']:
        n = int(r[ie] or 0); t = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip())
        op = t.split()[0].split('.')[0] if t else '?'
        c[op] += n; tot += n
    key = (k['name'][:50], tot)
    if key in seen or tot == 0: continue
    seen.add(key)
    uni = sum(v for o, v in c.items() if o.startswith('U') or o in ('S2UR', 'R2UR', 'LDCU'))
    br = sum(c[o] for o in ('BRA', 'BSSY', 'BSYNC', 'WARPSYNC', 'NOP', 'BREAK'))
    print("%s\n  warp-instructions %d; uniform-datapath %.1f%%; branch/reconvergence %.1f%%" % (k['name'][:70], tot, 100 * uni / tot, 100 * br / tot))
    print("  " + ", ".join("%s %.1f%%" % (o, 100 * v / tot) for o, v in c.most_common(22)))
