"""Developer aid: print the SASS of the hottest regions of an `ncu --page source --print-source sass --csv` dump
with executed-instruction counts (per window) and stall samples."""
import csv, sys
path = sys.argv[1]; nwin = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
lo = int(sys.argv[3]) if len(sys.argv) > 3 else None; hi = int(sys.argv[4]) if len(sys.argv) > 4 else None
rows = list(csv.reader(open(path)))[2:]
tot_i = sum(int(r[5]) for r in rows); tot_s = sum(int(r[4]) for r in rows)
print(f"total inst/window {tot_i/nwin:.0f} samples {tot_s}")
if lo is None:
    # hot regions: windows of 40 instructions ranked by samples
    W = 40; best = []
    for i in range(0, len(rows), W):
        s = sum(int(r[4]) for r in rows[i:i+W]); n = sum(int(r[5]) for r in rows[i:i+W])
        best.append((s, n, i))
    for s, n, i in sorted(best, reverse=True)[:25]:
        print(f"rows {i:5d}-{i+W:5d} samples {100*s/tot_s:5.1f}% inst {100*n/tot_i:5.1f}%")
else:
    for i in range(lo, hi):
        r = rows[i]
        print(f"{i:5d} {int(r[5])/nwin:9.1f} {100*int(r[4])/tot_s:5.2f}% thr {r[8]:>3s} {r[1]}")
