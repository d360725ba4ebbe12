"""Compact per-kernel table from an `ncu --set full` report (.ncu-rep) or its `--page raw --csv` export:
duration, DRAM bytes, DRAM / L2 / issue utilisation, occupancy, registers, shared-memory bank conflicts and
the top warp-stall reasons.  usage: python tools/ncu_kernels.py <report.ncu-rep | raw.csv> [> profiles/...txt]"""
import csv, io, subprocess, sys

path = sys.argv[1]
if path.endswith(".ncu-rep"):
    text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
else:
    text = open(path).read()
rows = list(csv.reader(io.StringIO(text)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]


def col(key):
    for i, n in enumerate(names):
        if n == key:
            return i
    return None


def val(r, key, default=float("nan")):
    i = col(key)
    if i is None or i >= len(r) or r[i] in ("", "n/a"):
        return default
    try:
        v = float(r[i].replace(",", ""))
    except ValueError:
        return default
    u = units[i]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
    return v * scale


stall_cols = [(i, n) for i, n in enumerate(names) if "smsp__average_warps_issue_stalled" in n and n.endswith("_per_issue_active.ratio")]
print("%-34s %8s %8s %8s %6s %6s %6s %6s %9s %5s %9s  %s" % ("kernel", "us", "rd MB", "wr MB", "dram%", "L2%", "issue%", "warps%", "winst", "regs", "bankconf", "top stalls (warps per issue)"))
for r in data:
    if len(r) < len(names) // 2:
        continue
    k = r[col("Kernel Name")].split("(")[0].replace("plh::", "")[:34]
    stalls = []
    for i, n in stall_cols:
        try:
            stalls.append((float(r[i].replace(",", "")), n.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
        except (ValueError, IndexError):
            pass
    stalls.sort(reverse=True)
    print("%-34s %8.2f %8.2f %8.2f %6.1f %6.1f %6.1f %6.1f %9d %5d %9d  %s" % (
        k, val(r, "gpu__time_duration.sum"), val(r, "dram__bytes_read.sum") / 1e6, val(r, "dram__bytes_write.sum") / 1e6,
        val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        val(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"), val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        int(val(r, "smsp__inst_executed.sum", 0)), int(val(r, "launch__registers_per_thread", 0)),
        int(val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 0)),
        ", ".join("%s %.1f" % (n, v) for v, n in stalls[:4])))
