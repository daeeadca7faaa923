"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
   python scripts/summarize_launches.py launches.csv [blocks] [title...]"""
import csv, re, sys
from collections import defaultdict

args = sys.argv[1:]
last = None
if "--last" in args:                      # only the last N launches of the list (e.g. one whole block)
    i = args.index("--last"); last = int(args[i + 1]); del args[i:i + 2]
path = args[0]
blocks = int(args[1]) if len(args) > 1 else 1
title = " ".join(args[2:])
tot = defaultdict(float); cnt = defaultdict(int)
rows = [l for l in open(path, errors="replace") if l.startswith('"')]
recs = [r for r in csv.DictReader(rows) if r.get("Metric Name") == "gpu__time_duration.sum"]
if last:
    recs = recs[-last:]
for r in recs:
    name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("apv::<unnamed>::", "").replace("void ", "")
    v = float(r["Metric Value"].replace(",", ""))
    if r.get("Metric Unit", "ns") in ("us", "usecond"):
        v *= 1e3
    elif r.get("Metric Unit") in ("ms", "msecond"):
        v *= 1e6
    tot[name] += v; cnt[name] += 1
total = sum(tot.values())
if title:
    print(title)
print(f"{blocks} block(s); cold-cache serialised times: compare SHARES")
print(f"total {total / 1e6:.1f} ms over {sum(cnt.values())} launches ({total / 1e6 / blocks:.1f} ms/block)\n")
print(f"{'kernel':52s} {'launches':>8s} {'total ms':>12s} {'share':>7s} {'ms/block':>12s}")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k[:52]:52s} {cnt[k]:8d} {tot[k] / 1e6:12.3f} {100 * tot[k] / total:6.1f}% {tot[k] / 1e6 / blocks:12.3f}")
