"""Per-kernel totals and shares from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X ...`).

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
iname, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= ival:
        continue
    name = re.sub(r"\(.*", "", r[iname]).replace("void ", "").replace("<unnamed>::", "")
    v = float(r[ival].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}[r[iunit]]
    tot[name][0] += 1
    tot[name][1] += v
total = sum(v[1] for v in tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'us/launch':>10s}")
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:60]:60s} {n:8d} {ms:10.3f} {ms / total:7.3f} {1e3 * ms / n:10.1f}")
print(f"{'total':60s} {sum(v[0] for v in tot.values()):8d} {total:10.3f}")
