"""Summarise `ncu --page source --print-source cuda,sass --csv` per CUDA source line (instructions, stall samples)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
data = []
for r in rows:
    if len(r) > 8 and r[0].isdigit():
        try:
            data.append((int(r[0]), r[1], int(r[4]), int(r[7]), int(r[8])))
        except ValueError:
            pass
tot_i = sum(d[3] for d in data)
tot_s = sum(d[2] for d in data)
print("total warp-inst %d, stall samples %d" % (tot_i, tot_s))
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print("%5.1f%% inst %5.1f%% stall  avg-thr %4.1f  L%-4d %s" % (100.0 * d[3] / max(1, tot_i), 100.0 * d[2] / max(1, tot_s),
                                                          d[4] / max(1, d[3]), d[0], d[1].strip()[:100]))
