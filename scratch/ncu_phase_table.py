"""Per-phase instruction / stall-sample table from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: python scratch/ncu_phase_table.py report.csv npoints"""
import csv
import sys

path, npts = sys.argv[1], float(sys.argv[2])
only = sys.argv[3] if len(sys.argv) > 3 else None          # keep only functions whose name contains this
rows = list(csv.reader(open(path)))
cur = None
keep = True
per_line = {}
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        keep = only is None or only in r[1]
        continue
    if r[0] == "Line No":
        hdr = {k: i for i, k in enumerate(r)}
        continue
    if hdr is None or r[2] != "-" or not keep:       # aggregated source-line rows carry "-" in the Address column
        continue
    try:
        ln = int(r[0])
        ins = int(r[hdr["Instructions Executed"]] or 0)
        smp = int(r[hdr["# Samples"]] or 0)
    except ValueError:
        continue
    key = (cur, ln)
    a = per_line.setdefault(key, [0, 0, r[1]])
    a[0] += ins
    a[1] += smp
tot_i = sum(v[0] for v in per_line.values())
tot_s = sum(v[1] for v in per_line.values())
print("total warp instructions %d (%.0f per point), samples %d" % (tot_i, tot_i / npts, tot_s))
by_file = {}
for (f, ln), v in per_line.items():
    b = by_file.setdefault(f, [0, 0])
    b[0] += v[0]; b[1] += v[1]
for f, b in by_file.items():
    print("  %-28s instr %5.1f%%  samples %5.1f%%" % (f, 100 * b[0] / tot_i, 100 * b[1] / tot_s))
print("hottest source lines by samples:")
for (f, ln), v in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:45]:
    print("  %5.1f%% smp %5.1f%% ins  %-24s:%4d  %s" % (100 * v[1] / tot_s, 100 * v[0] / tot_i, f, ln, v[2].strip()[:110]))

# ---- phases of pm_tc_kernel by source line ranges (markers looked up in the kernel source)
import os
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sea_ice_drift_b200", "csrc", "sid_pm_tc_kernel.cuh")).read().split("\n")


def find(s):
    for i, l in enumerate(src):
        if s in l:
            return i + 1
    return None


marks = [("helpers/ptx wrappers", 1), ("setup + point fetch", find("pm_tc_kernel(const PmArgs a")), ("1 stage window", find("---- 1. stage the window")),
         ("2 mma sums: squares LUT pass", find("---- 2 (tensor cores)")), ("2 mma sums: tiles (clear/gen/mma)", find("const int ntile_s =")),
         ("2 mma sums: sliding + den", find("// vertical sliding sums along the lane")), ("2 legacy sums", find("---- 2a.")),
         ("3 batch setup + gather", find("---- 3. angle batches")), ("mac: tile setup + clears", find("---- correlation on the tensor cores")),
         ("mac: row loop (gen+issue)", find("// Row loop.")), ("mac: commit/done wait", find("if (lane == 0 && wiw < g.nissue) tc_commit(&done_bar[wg]);")),
         ("epilogue (ld + normalise)", find("// epilogue: each warp group")), ("argmax merge / best angle", find("for (int a2 = 0; a2 < nb; ++a2)")),
         ("4 tail hand-off", find("---- 4. peak statistics")), ("end", 10 ** 9)]
marks = [m for m in marks if m[1] is not None]
common = {"normalisation (epilogue)": (218, 293), "window_den (2b)": (238, 249), "template sample (gather)": (294, 362), "keys / argmax": (19, 28), "tma/mbarrier": (486, 511)}
print("phase table (tc kernel file):")
for (name, start), (_, end) in zip(marks, marks[1:]):
    ins = sum(v[0] for (f, ln), v in per_line.items() if f == "sid_pm_tc_kernel.cuh" and start <= ln < end)
    smp = sum(v[1] for (f, ln), v in per_line.items() if f == "sid_pm_tc_kernel.cuh" and start <= ln < end)
    print("  %-30s instr %5.1f%% (%6.0f / point)   samples %5.1f%%" % (name, 100 * ins / tot_i, ins / npts, 100 * smp / tot_s))
print("shared device functions (sid_common.cuh):")
for name, (lo, hi) in common.items():
    ins = sum(v[0] for (f, ln), v in per_line.items() if f == "sid_common.cuh" and lo <= ln < hi)
    smp = sum(v[1] for (f, ln), v in per_line.items() if f == "sid_common.cuh" and lo <= ln < hi)
    print("  %-30s instr %5.1f%% (%6.0f / point)   samples %5.1f%%" % (name, 100 * ins / tot_i, ins / npts, 100 * smp / tot_s))
