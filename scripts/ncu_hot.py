"""Summarise an ncu --page source --csv export: stall reasons and the hottest SASS instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
def fl(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return 0.0
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != 'Address']
si, ii, src = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
tot = sum(fl(r[si]) for r in data)
tinst = sum(fl(r[ii]) for r in data)
print('total samples', tot, 'instr rows', len(data), 'warp instr executed', tinst)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(fl(r[hdr.index(h)]) for r in data) for h in stalls}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f'{k:28s} {v:10.0f} {100*v/tot:5.1f}%')
print('--- top instructions by samples')
for r in sorted(data, key=lambda r: -fl(r[si]))[:topn]:
    best = max(stalls, key=lambda h: fl(r[hdr.index(h)]))
    print(f'{r[0][-5:]:>5s} {fl(r[si]):8.0f} {100*fl(r[si])/tot:5.1f}% exec={fl(r[ii]):>11.0f} {best:20s} {r[src][:90]}')
