'''Aggregate the warp-stall samples of an ncu report per kernel stage (source-line ranges marked by comments).
usage: python scripts/ncu_stages.py report.ncu-rep source.cu marker1 marker2 ...'''
import csv, subprocess, sys, io
rep, srcfile = sys.argv[1], sys.argv[2]
markers = sys.argv[3:]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Line No')
hdr = rows[h]
iS = hdr.index('# Samples')
iI = hdr.index('Instructions Executed')
stall_cols = [(i, x) for i, x in enumerate(hdr) if x.startswith('stall_') and 'Not' not in x]
lines = {}
sass = []
cur = None
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    if r[0]:
        try:
            cur = int(r[0]); lines[cur] = (int(r[iS]), r)
        except ValueError:
            pass
    elif cur is not None:
        try:
            sass.append((cur, r[3].strip(), int(r[iS]), int(r[iI]), r))
        except ValueError:
            pass
tot = sum(s for s, _ in lines.values())
src = open(srcfile).read().split('\n')
def find(txt):
    return next(i + 1 for i, l in enumerate(src) if txt in l)
marks = [('head', 1)] + [(m, find(m)) for m in markers] + [('end', len(src) + 1)]
print('total samples', tot)
for (name, a), (_, b) in zip(marks[:-1], marks[1:]):
    ss = sum(s for ln, (s, _) in lines.items() if a <= ln < b)
    st = {}
    for ln, (s, r) in lines.items():
        if a <= ln < b:
            for i, x in stall_cols:
                if r[i].isdigit():
                    st[x] = st.get(x, 0) + int(r[i])
    ninst = sum(n for ln, _, _, n, _ in sass if a <= ln < b)
    nfp64 = sum(n for ln, op, _, n, _ in sass if a <= ln < b and op.split()[0].lstrip('@!P0123456789 ').startswith(('DFMA', 'DMUL', 'DADD')))
    top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
    print('%-28s %4d-%4d %7d %5.1f%%  inst %.3g fp64 %.3g | ' % (name[:28], a, b - 1, ss, 100 * ss / tot, ninst, nfp64), ', '.join('%s %.0f%%' % (k[6:], 100 * v / max(ss, 1)) for k, v in top))
if '--lines' in sys.argv:
    pass
print()
for ln, (s, r) in sorted(lines.items()):
    if s > tot * 0.012:
        print(ln, s, '%.1f%%' % (100 * s / tot), src[ln - 1].strip()[:110])
