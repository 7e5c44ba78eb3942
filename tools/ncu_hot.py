"""Hot instructions of the first kernel in an `ncu --page source --csv` export: python tools/ncu_hot.py src.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
heads = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdr = rows[heads[0]]
end = heads[1] - 1 if len(heads) > 1 else len(rows)
data = [r for r in rows[heads[0] + 1:end] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
num = lambda r, h: int(float(r[ix[h]] or 0))
tot = sum(num(r, '# Samples') for r in data)
print('instructions', len(data), 'samples', tot, 'warp-instr executed', sum(num(r, 'Instructions Executed') for r in data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
s = collections.Counter()
for r in data:
    for h in stalls:
        s[h] += num(r, h)
print(s.most_common(8))
top = sorted(range(len(data)), key=lambda k: -num(data[k], '# Samples'))[:top_n]
for k in sorted(top):
    r = data[k]
    st = sorted(((num(r, h), h[6:]) for h in stalls), reverse=True)[:2]
    print(k, num(r, '# Samples'), num(r, 'Instructions Executed'), r[ix['Source']][:64], st)
