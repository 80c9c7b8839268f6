"""Per-kernel shares of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file f.csv ...`).
usage: python tools/launch_summary.py f.csv "header comment" > profiles/launches_summary.csv"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if len(r) <= iv or r[ik] == "Kernel Name":
        continue
    tot[r[ik]] += float(r[iv].replace(",", "")) * scale[r[iu]]
    cnt[r[ik]] += 1
allus = sum(tot.values())
for c in sys.argv[2:]:
    print("# " + c)
print("kernel,launches,total_us,share")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print('"%s",%d,%.1f,%.4f' % (k, cnt[k], v, v / allus))
