"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
Usage: python scripts/summarise_launches.py launches.csv n_blocks "command line" > summary.txt"""
import csv, io, re, sys
from collections import defaultdict

path, nblk = sys.argv[1], int(sys.argv[2])
cmd = sys.argv[3] if len(sys.argv) > 3 else ""
lines = open(path, errors="replace").read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
    tot[name] += ms; cnt[name] += 1
total = sum(tot.values())
print(cmd)
print("%d blocks; cold-cache serialised times: compare SHARES" % nblk)
print("total %.1f ms over %d launches (%.1f ms/block)\n" % (total, sum(cnt.values()), total / nblk))
print("%-52s %8s %12s %7s %12s" % ("kernel", "launches", "total ms", "share", "ms/block"))
for k in sorted(tot, key=lambda k: -tot[k]):
    print("%-52s %8d %12.3f %6.1f%% %12.3f" % (k[:52], cnt[k], tot[k], 100 * tot[k] / total, tot[k] / nblk))
