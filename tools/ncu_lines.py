"""Warp-stall samples per CUDA source line of one kernel.
usage: ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME --launch-count 1 > f.csv; python tools/ncu_lines.py f.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
cur = None
out = []
for r in rows:
    if r and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
    if len(r) > 8 and r[0].isdigit():
        try:
            out.append((int(r[6]), cur, int(r[0]), r[1].strip()[:120]))
        except ValueError:
            pass
tot = sum(o[0] for o in out)
print("samples", tot)
for o in sorted(out, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%7d %5.1f%% %s:%d  %s" % (o[0], 100.0 * o[0] / max(tot, 1), o[1], o[2], o[3]))
