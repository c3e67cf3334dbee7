"""Aggregate an ncu source-page export per CUDA source line (development tool).
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv ; python tools/ncu_lines.py X.csv [top]
Prints the lines with the most executed warp instructions and stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
agg = {}
fname = None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    if r[0]:  # a source line starts a group; following rows with empty line no are its SASS
        cur = (fname, int(r[0]), r[1].strip()[:90])
        agg.setdefault(cur, [0, 0])
    try:
        agg[cur][0] += int(r[iI] or 0)
        agg[cur][1] += int(r[iS] or 0)
    except (ValueError, NameError):
        pass
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[1] for v in agg.values()) or 1
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][int(len(sys.argv) > 3)])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * v[0] / tot_i, 100.0 * v[1] / tot_s, k[0], k[1], k[2]))
