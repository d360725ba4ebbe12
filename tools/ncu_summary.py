"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / avg, and the
share of ONE head step (every libplhead kernel launches exactly once per step, so the per-step share is
avg / sum of avgs; the raw time share also counts bench.py's roofline pass, which relaunches the main
pass alone).  ncu times are cold-cache and serialised: compare shares, not absolutes."""
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
ours = {k: v for k, v in agg.items() if "plh::" in k or k.startswith("decode_") or "loss_main" in k or "ohem" in k or "score_keys" in k}
step = sum(sum(v) / len(v) for v in ours.values())
print("%-50s %5s %9s %11s" % ("kernel", "n", "avg us", "step share"))
for k, v in agg.items():
    avg = sum(v) / len(v)
    print("%-50s %5d %9.2f %10s" % (k, len(v), avg / 1e3, ("%.1f%%" % (100 * avg / step)) if k in ours else "-"))
print("sum of per-kernel averages (one step, serialised, cold): %.1f us" % (step / 1e3))
