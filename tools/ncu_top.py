#!/usr/bin/env python
"""Top stalled SASS instructions per kernel of an .ncu-rep (source page), for reading a profile on a box without a GPU."""
import csv, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; kern.append(cur); continue
    if r and r[0] == 'Address': cur['hdr'] = r; continue
    if cur is not None and len(r) > 5: cur['rows'].append(r)
k = kern[which]; h = k['hdr']
si = h.index('# Samples'); ex = h.index('Instructions Executed')
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(r[si]) for r in k['rows'])
print(k['name'], 'total samples', tot, 'instructions', len(k['rows']))
agg = {h[i]: sum(int(r[i]) for r in k['rows']) for i in stall_cols}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
# cumulative samples by instruction-index buckets of 250 (shows which code region = role is hot)
bucket = {}
for i, r in enumerate(k['rows']):
    bucket[i // 250] = bucket.get(i // 250, 0) + int(r[si])
print('samples per 250-instruction bucket:', bucket)
top = sorted(range(len(k['rows'])), key=lambda i: -int(k['rows'][i][si]))[:topn]
for i in sorted(top):
    r = k['rows'][i]
    st = sorted(((int(r[j]), h[j]) for j in stall_cols), reverse=True)[:2]
    print(i, r[1].strip()[:64].ljust(64), r[si].rjust(6), r[ex].rjust(8), st)
