"""Top sampled SASS instructions per kernel from `ncu -i REP --page source --csv` output (stdin or file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
ks, cur, seen = [], None, set()
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = []
        if r[1] not in seen: ks.append((r[1], cur)); seen.add(r[1])
        continue
    if cur is not None and len(r) == len(hdr) and r[iS].isdigit(): cur.append(r)
for name, data in ks:
    tot = sum(int(r[iS]) for r in data)
    print('==', name[:90], 'instr', len(data), 'samples', tot)
    agg = {}
    for r in data:
        for i in stall: agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(r[i])
    print('   stalls:', sorted(agg.items(), key=lambda x: -x[1])[:8])
    top = sorted(range(len(data)), key=lambda k: -int(data[k][iS]))[:ntop]
    for k in sorted(top):
        r = data[k]
        st = sorted(((hdr[i][6:], int(r[i])) for i in stall if int(r[i]) > 0), key=lambda x: -x[1])[:2]
        print(f'   {k:5d} {int(r[iS]):6d} {int(r[iEx]):9d} {r[iSrc].strip()[:64]:64s} {st}')
