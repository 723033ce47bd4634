"""Per-CUDA-source-line summary of `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out, cur_file, hdr = [], None, None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    elif r and r[0] == 'Line No':
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] not in ('', 'Line No'):
        out.append((cur_file, r))
def fl(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return 0.0
si, ii, ti = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
ts = sum(fl(r[si]) for _, r in out)
tinst = sum(fl(r[ii]) for _, r in out)
tthr = sum(fl(r[ti]) for _, r in out)
print('samples', ts, 'warp instr', tinst, 'avg active lanes', tthr / max(tinst, 1))
for f, r in sorted(out, key=lambda fr: -fl(fr[1][si]))[:topn]:
    print(f'{f[:14]:14s}:{r[0]:>4s} smp {100*fl(r[si])/ts:5.1f}% inst {100*fl(r[ii])/tinst:5.1f}% lanes {fl(r[ti])/max(fl(r[ii]),1):4.1f}  {r[1].strip()[:100]}')
if len(sys.argv) > 3:
    # region shares: "name:lo-hi,name:lo-hi" on the first file
    for spec in sys.argv[3].split(','):
        name, rng = spec.split(':')
        lo, hi = map(int, rng.split('-'))
        sel = [r for f, r in out if f.startswith('himm_kernels') and lo <= int(r[0]) <= hi]
        print(f'{name:12s} smp {100*sum(fl(r[si]) for r in sel)/ts:5.1f}% inst {100*sum(fl(r[ii]) for r in sel)/tinst:5.1f}%')
    other = [r for f, r in out if not f.startswith('himm_kernels')]
    print(f'{"other files":12s} smp {100*sum(fl(r[si]) for r in other)/ts:5.1f}% inst {100*sum(fl(r[ii]) for r in other)/tinst:5.1f}%')
