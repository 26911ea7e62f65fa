"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name: count, total time, share."""
import collections, csv, sys
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
    if not hdr:
        print(path, "no launches"); continue
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr[0] + 1:]:
        if len(r) < 10: continue
        name = r[4].split("(")[0].replace("void ", "")[:64]
        try: agg[name][1] += float(r[-1].replace(",", "")) / 1e3
        except ValueError: continue
        agg[name][0] += 1
    tot = sum(v[1] for v in agg.values())
    print(f"== {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"   {k:64s} n={v[0]:4d} {v[1]:10.1f} us {100 * v[1] / tot:5.1f}%")
