#!/usr/bin/env python
"""Copy the judged evidence of one GPU round from gpurun_out/<tag>/ into profiles/ (tracked):
launch list (raw csv + per-kernel aggregate), ncu --set full summaries, bench JSON lines.
usage: python scripts/make_profiles.py <tag> <round-prefix>     e.g.  r01p r01"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launch_table(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["%-48s %7s %12s %7s" % ("kernel", "launches", "total ms", "share")]
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-48s %7d %12.3f %6.1f%%" % (n[:48], c, ms, 100 * ms / tot))
    out.append("%-48s %7s %12.3f" % ("total", "", tot))
    step = ("k_integrate", "k_fourier_table", "k_postprocess", "k_select", "k_seed", "k_coef_from_philox", "k_gather",
            "k_wave_stats", "k_assign_slots", "k_scan_counts")
    sub = {n: v for n, v in agg.items() if any(k in n for k in step)}
    st = sum(v[1] for v in sub.values())
    out.append("")
    out.append("# kernels of the timed step only (what bench.py's kernel_share_of_step covers)")
    for n, (c, ms) in sorted(sub.items(), key=lambda kv: -kv[1][1]):
        out.append("%-48s %7d %12.3f %6.1f%%" % (n[:48], c, ms, 100 * ms / st))
    return "\n".join(out)


def main():
    tag, pre = sys.argv[1], sys.argv[2]
    src = os.path.join(ROOT, "gpurun_out", tag)
    dst = os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    for name in ("bench.json", "bench_reference.json", "gpu.txt", "pytest_gpu.log", "smoke.log"):
        if os.path.exists(os.path.join(src, name)):
            shutil.copy(os.path.join(src, name), os.path.join(dst, "%s_%s" % (pre, name)))
    lc = os.path.join(src, "launches.csv")
    if os.path.exists(lc):
        shutil.copy(lc, os.path.join(dst, pre + "_launches.csv"))
        with open(os.path.join(dst, pre + "_launches_by_kernel.txt"), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 3 --no-cpu --interp-queries 4194304\n")
            f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's kernel_share_of_step\n")
            f.write(launch_table(lc) + "\n")
    for rep, args in (("prof_integrate", []), ("prof_interp", []), ("prof_poi", []), ("prof_windstats", []), ("prof_thermo", []), ("prof_others", None)):
        path = os.path.join(src, rep + ".ncu-rep")
        pre_made = os.path.join(src, rep + "_summary.txt")         # summarised on the GPU box (gpu_round.sh)
        pre_raw = os.path.join(src, rep + "_raw.csv")
        if not (os.path.exists(path) or os.path.exists(pre_made) or os.path.exists(pre_raw)):
            continue
        if args is not None:
            if os.path.exists(pre_made):
                txt = open(pre_made).read()
            else:
                txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), path, "40"],
                                     capture_output=True, text=True).stdout
            open(os.path.join(dst, "%s_%s_summary.txt" % (pre, rep)), "w").write(
                "# ncu --set full --clock-control none --import-source on ; summarised by scripts/ncu_summary.py\n" + txt)
        else:
            # several kernels in one report: one raw-metric block each
            if os.path.exists(pre_raw):
                raw = open(pre_raw).read()
            else:
                raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(raw.splitlines()))
            hdr, units = rows[0], rows[1]
            keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
                    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
                    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
                    "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
                    "smsp__thread_inst_executed_per_inst_executed.ratio"]
            with open(os.path.join(dst, "%s_%s_summary.txt" % (pre, rep)), "w") as f:
                f.write("# ncu --set full --clock-control none: the small kernels of one wave\n")
                for r in rows[2:]:
                    for k in keys:
                        if k in hdr:
                            i = hdr.index(k)
                            f.write("%-66s %-14s %s\n" % (k, units[i], r[i]))
                    f.write("\n")
    print(os.listdir(dst))


if __name__ == "__main__":
    main()
