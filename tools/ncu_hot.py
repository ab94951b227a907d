"""Top SASS instructions of an `ncu --page source --csv` dump by stall samples."""
import csv
import sys

r = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = r[1]
idx = {n: i for i, n in enumerate(h)}
rows = []
for k, row in enumerate(r[2:]):
    try:
        v = int(row[idx["# Samples"]])
    except (ValueError, IndexError):
        continue
    rows.append((v, k, row[idx["Source"]].strip(), row[idx["Instructions Executed"]],
                 row[idx.get("L1 Tag Requests Global", 0)], row[idx.get("L1 Wavefronts Shared", 0)]))
tot = sum(v for v, *_ in rows)
print("total samples", tot, "instructions", len(rows))
for v, k, s, n, tg, ws in sorted(rows, reverse=True)[:top]:
    print("%6d %5.1f%%  #%-5d exec=%-9s tag=%-9s shw=%-9s %s" % (v, 100.0 * v / tot, k, n, tg, ws, s[:90]))
