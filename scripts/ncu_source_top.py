#!/usr/bin/env python
"""Top stall locations of one kernel instance from `ncu --page source --csv` output.
usage: ncu -i rep --page source --csv --kernel-name K > src.csv ; python scripts/ncu_source_top.py src.csv [instance] [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
inst = int(sys.argv[2]) if len(sys.argv) > 2 else -1
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
starts.append(len(rows))
k = inst if inst >= 0 else len(starts) - 2
blk = rows[starts[k]:starts[k + 1]]
hdr = blk[1]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in blk[2:] if len(r) == len(hdr)]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[col['# Samples']] or 0) for r in body)
print(blk[0][1][:80], 'instance', k, 'total samples', tot)
agg = {s: sum(int(r[col[s]] or 0) for r in body) for s in stalls}
print('by reason:', ', '.join('%s %.1f%%' % (s[6:], 100.0 * v / max(tot, 1)) for s, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 50 > tot))
body.sort(key=lambda r: -int(r[col['# Samples']] or 0))
for r in body[:top]:
    n = int(r[col['# Samples']] or 0)
    why = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print('%5.1f%%  %-70s %s' % (100.0 * n / max(tot, 1), r[col['Source']][:70], ' '.join('%s:%d' % (b, a) for a, b in why if a)))
