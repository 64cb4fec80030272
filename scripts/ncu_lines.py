#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares from `ncu -i rep --page source --csv --print-source cuda,sass`.
usage: python scripts/ncu_lines.py src.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == 'Line No')
col = {}
for i, h in enumerate(hdr):
    col.setdefault(h, i)
ie, sm = col['Instructions Executed'], col['# Samples']
lines = collections.OrderedDict()
cur_file = ''
for r in rows:
    if r and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    if len(r) < len(hdr) - 3 or r[0] == 'Line No':
        continue
    if r[2] == '-':   # a source line with aggregated metrics
        key = (cur_file, r[0])
        try:
            lines[key] = [r[1], int(r[ie] or 0), int(r[sm] or 0)]
        except ValueError:
            pass
tot_i = sum(v[1] for v in lines.values()) or 1
tot_s = sum(v[2] for v in lines.values()) or 1
print('total warp instructions', tot_i, 'samples', tot_s)
for (f, ln), (src, i, s) in sorted(lines.items(), key=lambda kv: -kv[1][2])[:top]:
    print('%5.1f%% smp %5.1f%% inst  %s:%s  %s' % (100.0 * s / tot_s, 100.0 * i / tot_i, f, ln, src.strip()[:100]))
