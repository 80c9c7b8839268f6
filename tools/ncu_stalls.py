"""Aggregate warp-stall samples of one kernel from an ncu source-page CSV.
usage: ncu -i rep --page source --csv --kernel-name regex:NAME --launch-count 1 > f.csv; python tools/ncu_stalls.py f.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
idx = {k: i for i, k in enumerate(hdr)}
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot = {s: 0 for s in stalls}
samples = 0
data = []


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


for r in rows[h + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    n = num(r[idx["# Samples"]])
    samples += n
    st = {s: num(r[idx[s]]) for s in stalls}
    for s in stalls:
        tot[s] += st[s]
    data.append((n, r[idx["Source"]].strip(), {s: v for s, v in st.items() if v > 0}))
print("total samples", samples)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    print("  %-24s %7d  %.1f%%" % (s, v, 100.0 * v / max(samples, 1)))
print("top instructions:")
for n, src, st in sorted(data, key=lambda d: -d[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print("  %5d  %-64s %s" % (n, src[:64], st))
