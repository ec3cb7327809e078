"""Dynamic SASS instruction mix of one kernel from an ncu source-page CSV
(ncu -i rep --page source --csv --kernel-id ...):  opcode -> executed warp instructions,
plus the hottest instructions by stall samples.  usage: sass_hot.py file.csv [cells]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
cells = float(sys.argv[2]) if len(sys.argv) > 2 else 4096.0 * 4096
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]
iS, iX, iT, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
mix = collections.Counter(); tot = 0; samples = collections.Counter()
body = []
for r in rows[hdr_i + 1:]:
    if r and r[0] in ('Kernel Name', 'Address'): break      # first launch only
    if len(r) <= iT: continue
    src = r[iS].strip()
    src2 = re.sub(r'^@!?U?P\w+\s+', '', src)
    op = src2.split()[0].split('.')[0] if src2 else '?'
    n = int(r[iX] or 0)
    mix[op] += n; tot += n
    samples[op] += int(r[iSm] or 0)
    body.append((n, int(r[iSm] or 0), src))
print('warp instructions: %d   per cell-lane: %.1f' % (tot, tot * 32 / cells))
ts = sum(samples.values())
for op, n in mix.most_common(28):
    print('  %-10s %6.2f %%  (%.1f / cell)   stall samples %5.1f %%' % (op, 100.0 * n / tot, n * 32 / cells, 100.0 * samples[op] / max(ts, 1)))
if '--top' in sys.argv:
    for n, s, src in sorted(body, key=lambda t: -t[1])[:40]:
        print('%8d %6d  %s' % (n, s, src))
