"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr_i + 1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"<.*", "", r[kn])[:70]
    agg[name][0] += 1
    agg[name][1] += float(r[mv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print("total launches %d, total kernel time %.3f ms" % (sum(v[0] for v in agg.values()), tot / 1e6))
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%6.2f%%  %9.3f ms  x%-5d %s" % (100 * t / tot, t / 1e6, n, name))
