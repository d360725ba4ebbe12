"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / avg / share."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        agg.setdefault(d["Kernel Name"].split("(")[0][:48], []).append(float(d["Metric Value"].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print("%-50s %5s %9s %7s" % ("kernel", "n", "avg us", "share"))
for k, v in agg.items():
    print("%-50s %5d %9.2f %6.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
