"""Hottest source lines (instructions executed / warp samples) of an `ncu --page source --csv --print-source cuda,sass` dump.
usage: python scratch/ncu_lines.py report.csv npoints [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
npts = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
hdr = None; cur = None; per = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = {k: i for i, k in enumerate(r)}; continue
    if hdr is None or len(r) < 3 or r[2] != "-": continue
    try:
        ln = int(r[0]); ins = int(r[hdr["Instructions Executed"]] or 0); smp = int(r[hdr["# Samples"]] or 0)
    except ValueError: continue
    a = per.setdefault((cur, ln), [0, 0, r[1]]); a[0] += ins; a[1] += smp
ti = sum(v[0] for v in per.values()); ts = sum(v[1] for v in per.values())
print("total warp instructions %d (%.0f per point), samples %d" % (ti, ti / npts, ts))
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% ins (%6.0f/pt) %5.1f%% smp  %s:%d  %s" % (100 * v[0] / ti, v[0] / npts, 100 * v[1] / ts, k[0], k[1], v[2].strip()[:100]))
