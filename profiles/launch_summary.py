"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hd = rows[h]; kn = hd.index("Kernel Name"); mv = hd.index("Metric Value"); mu = hd.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    if r[mu] in ("ns", "nsecond"): v /= 1e3
    elif r[mu] in ("ms", "msecond"): v *= 1e3
    k = re.sub(r"\(.*", "", r[kn])[:100]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("total %.1f us over %d launches" % (tot, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%10.1f us %5d %5.1f%% %s" % (v[1], v[0], 100 * v[1] / tot, k))
