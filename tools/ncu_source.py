"""Hottest source lines of one kernel in an ncu report (needs -lineinfo and --import-source on).
usage: python tools/ncu_source.py <report.ncu-rep> <kernel-name-regex> [top N] [launch-skip]"""
import csv, io, os, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                       "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
fname, names, data = "", None, []
for r in csv.reader(io.StringIO(text)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1])
    elif r[0] == "Line No":
        names = r
        isamp, iinst = names.index("# Samples"), names.index("Instructions Executed")
    elif names and r[0].isdigit() and len(r) > iinst:
        try:
            data.append((int(r[isamp] or 0), int(r[iinst] or 0), fname, r[0], r[1].strip()))
        except ValueError:
            pass
tot = sum(d[0] for d in data) or 1
toti = sum(d[1] for d in data) or 1
print("kernel %s: %d samples, %d warp instructions" % (kern, tot, toti))
for s, n, f, line, src in sorted(data, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%-4s %s" % (100.0 * s / tot, 100.0 * n / toti, f, line, src[:120]))
