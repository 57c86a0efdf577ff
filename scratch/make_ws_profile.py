"""Text summary of one `ncu --set full --import-source on` capture of a step's two kernels (pm_ws_kernel + pm_tail_kernel):
headline metrics, instructions / samples per role section, hottest lines, shared-memory bank conflicts per line.
usage: python scratch/make_ws_profile.py gpurun_out/ws_final.ncu-rep 39425 > profiles/r02_pm_ws_ncu_roles.txt"""
import csv, io, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, npts = sys.argv[1], float(sys.argv[2])


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
rows = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units = rows[0], rows[1]
print("ncu --set full --clock-control none --import-source on, cfg2 (39 425 grid points, 3 angles, img_size 35, border 20), one launch each")
print("(per-launch times under ncu are cold-cache and serialised; bench.py's CUDA-event times are the values to quote)\n")
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print("== " + d["Kernel Name"])
    for k in WANT:
        if k in d and d[k] not in ("", None):
            print("  %-92s %16s %s" % (k, d[k], u[k]))
    print()


def per_line(kernel):
    txt = ncu("--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kernel)
    hdr = None; cur = None; per = {}
    for r in csv.reader(io.StringIO(txt)):
        if not r: continue
        if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
        if r[0] == "Line No": hdr = {k: i for i, k in enumerate(r)}; continue
        if hdr is None or len(r) < 3 or r[2] != "-": continue
        try: ln = int(r[0])
        except ValueError: continue
        def g(k):
            try: return int(r[hdr[k]] or 0)
            except (ValueError, KeyError): return 0
        a = per.setdefault((cur, ln), [0, 0, 0, 0, r[1]])
        a[0] += g("Instructions Executed"); a[1] += g("# Samples"); a[2] += g("L1 Wavefronts Shared"); a[3] += g("L1 Wavefronts Shared Excessive")
    return per


def sections(fname):
    src = open(os.path.join(ROOT, "sea_ice_drift_b200/csrc", fname)).read().split("\n")
    marks = []
    for i, l in enumerate(src, 1):
        m = re.search(r"// (=====+ |---- )(.*)", l)
        if m: marks.append((i, m.group(2)[:56]))
    return [(1, "helpers above the kernel")] + marks + [(len(src) + 1, "end")]


def report(kernel, fname, top=24):
    per = per_line(kernel)
    ti = sum(v[0] for v in per.values()); ts = sum(v[1] for v in per.values()) or 1
    tw = sum(v[2] for v in per.values()) or 1; te = sum(v[3] for v in per.values()) or 1
    print("== %s: %d warp instructions (%.0f per grid point), %d warp samples; shared-memory wavefronts %d, of them %d (%.0f %%) excess (bank conflicts)"
          % (kernel, ti, ti / npts, ts, tw, te, 100.0 * te / tw))
    print("-- by section of %s (markers in the source):" % fname)
    marks = sections(fname)
    for (lo, name), (hi, _) in zip(marks[:-1], marks[1:]):
        sel = [v for k, v in per.items() if k[0] == fname and lo <= k[1] < hi]
        i = sum(v[0] for v in sel); s = sum(v[1] for v in sel)
        if i: print("   %4d-%4d %-58s %7.0f instr/pt %5.1f %% samples" % (lo, hi - 1, name, i / npts, 100.0 * s / ts))
    others = {}
    for k, v in per.items():
        if k[0] != fname: o = others.setdefault(k[0], [0, 0]); o[0] += v[0]; o[1] += v[1]
    for f, v in sorted(others.items(), key=lambda kv: -kv[1][0]):
        if v[0] / npts >= 20: print("   %-68s %7.0f instr/pt %5.1f %% samples" % ("(inlined from) " + f, v[0] / npts, 100.0 * v[1] / ts))
    print("-- hottest lines:")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print("   %5.1f %% ins (%5.0f/pt) %5.1f %% smp  %s:%d  %s" % (100.0 * v[0] / ti, v[0] / npts, 100.0 * v[1] / ts, k[0], k[1], v[4].strip()[:96]))
    print("-- shared-memory bank conflicts (excess wavefronts) by line:")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][3])[:10]:
        if v[3]: print("   %5.1f %% of the excess, %5.1f %% of all wavefronts  %s:%d  %s" % (100.0 * v[3] / te, 100.0 * v[2] / tw, k[0], k[1], v[4].strip()[:96]))
    print()


report("pm_ws_kernel", "sid_pm_ws_kernel.cuh")
report("pm_tail_kernel", "sid_pm_kernel.cuh", top=16)
