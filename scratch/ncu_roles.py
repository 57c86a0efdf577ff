"""Instructions / samples of pm_ws_kernel per role section (line ranges from the section markers in the source).
usage: python scratch/ncu_roles.py report_source.csv npoints"""
import csv, re, sys, os
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sea_ice_drift_b200/csrc/sid_pm_ws_kernel.cuh")).read().split("\n")
marks = []
for i, l in enumerate(src, 1):
    m = re.search(r"// (=====+ |---- )(.*)", l)
    if m: marks.append((i, m.group(2)[:44]))
marks = [(1, "helpers (ws_wait, exact pass fn)")] + marks + [(len(src) + 1, "end")]
rows = list(csv.reader(open(sys.argv[1]))); npts = float(sys.argv[2])
hdr = None; cur = None; per = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = {k: i for i, k in enumerate(r)}; continue
    if hdr is None or len(r) < 3 or r[2] != "-": continue
    try: ln = int(r[0]); ins = int(r[hdr["Instructions Executed"]] or 0); smp = int(r[hdr["# Samples"]] or 0)
    except ValueError: continue
    a = per.setdefault((cur, ln), [0, 0]); a[0] += ins; a[1] += smp
F = "sid_pm_ws_kernel.cuh"
ts = sum(v[1] for v in per.values())
for (lo, name), (hi, _) in zip(marks[:-1], marks[1:]):
    i = sum(v[0] for k, v in per.items() if k[0] == F and lo <= k[1] < hi); s = sum(v[1] for k, v in per.items() if k[0] == F and lo <= k[1] < hi)
    print("%4d-%4d %-46s %7.0f instr/pt %5.1f%% samples" % (lo, hi - 1, name, i / npts, 100.0 * s / ts))
others = {}
for k, v in per.items():
    if k[0] != F: o = others.setdefault(k[0], [0, 0]); o[0] += v[0]; o[1] += v[1]
for f, v in others.items(): print("%-56s %7.0f instr/pt %5.1f%% samples" % (f, v[0] / npts, 100.0 * v[1] / ts))
