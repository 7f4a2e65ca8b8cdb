"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
   python tools/ncu_source_lines.py dump.csv [top]  ->  samples / executed warp-instructions per file:line"""
import csv, sys, collections, os
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.defaultdict(lambda: [0, 0, ""]); cur_file = None; hdr = None; launches = 0
for row in csv.reader(open(path, newline="")):
    if not row: continue
    if row[0] == "File Path": cur_file = os.path.basename(row[1]); continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No":
        hdr = {k: i for i, k in enumerate(row)}; hdr_s = row.index("# Samples"); hdr_i = row.index("Instructions Executed"); continue
    if hdr is None or row[0] == "": continue
    try: ln = int(row[0])
    except ValueError: continue
    a = agg[(cur_file, ln)]
    a[0] += int(row[hdr_s] or 0); a[1] += int(row[hdr_i] or 0); a[2] = row[1].strip()[:90]
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print("total samples %d, warp-instructions %d" % (tot_s, tot_i))
byfile = collections.defaultdict(lambda: [0, 0])
for (f, l), a in agg.items(): byfile[f][0] += a[0]; byfile[f][1] += a[1]
for f, a in sorted(byfile.items(), key=lambda x: -x[1][0]): print("%-28s samples %6.2f%%  inst %6.2f%%" % (f, 100 * a[0] / max(tot_s, 1), 100 * a[1] / max(tot_i, 1)))
print()
for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%-22s:%4d  smp %5.2f%%  inst %5.2f%%  cyc/inst %5.1f | %s" % (f, l, 100 * a[0] / tot_s, 100 * a[1] / tot_i, (a[0] / tot_s) / max(a[1] / tot_i, 1e-9), a[2]))
if len(sys.argv) > 3:   # ranges file: "name file lo hi" per line
    print()
    for line in open(sys.argv[3]):
        name, f, lo, hi = line.split(); lo, hi = int(lo), int(hi)
        s = sum(a[0] for (ff, l), a in agg.items() if ff == f and lo <= l <= hi); i = sum(a[1] for (ff, l), a in agg.items() if ff == f and lo <= l <= hi)
        print("%-18s samples %6.2f%%  inst %6.2f%%" % (name, 100 * s / tot_s, 100 * i / tot_i))
