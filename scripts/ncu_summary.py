#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key raw metrics, stall mix, opcode mix, hottest source lines.
usage: python scripts/ncu_summary.py <report.ncu-rep> [n_lines]"""
import collections
import csv
import io
import subprocess
import sys


def ncu(rep, *args):
    extra = ["--launch-skip", sys.argv[3], "--launch-count", "1"] if len(sys.argv) > 3 else []
    return subprocess.run(["ncu", "-i", rep] + extra + list(args), capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    n_lines = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active", "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
    print("== raw metrics")
    for k in keys:
        for i, h in enumerate(hdr):
            if h == k:
                print("  %-70s %-14s %s" % (h, units[i], vals[i]))
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    h_at = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h_at]
    rows = rows[h_at - 1:]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    ns = ninst = nthr = 0
    opc = collections.defaultdict(lambda: [0, 0])
    n_kernels = 0
    for r in rows:
        if r and r[0] == "Kernel Name":
            n_kernels += 1
            continue
        if n_kernels > 1:
            break                       # first selected launch only
        if len(r) < len(hdr) or r[0] == "Address":
            continue
        s = int(r[ix["# Samples"]] or 0)
        ie = int(r[ix["Instructions Executed"]] or 0)
        ns += s; ninst += ie; nthr += int(r[ix["Thread Instructions Executed"]] or 0)
        for h in stall_cols:
            tot[h] += int(r[ix[h]] or 0)
        src = r[ix["Source"]].split()
        op = src[0] if src else ""
        if op.startswith("@") and len(src) > 1:
            op = src[1]
        op = op.split(".")[0]
        opc[op][0] += ie; opc[op][1] += s
    print("== SASS: %d instructions, %d samples, %.3e warp-inst executed, %.2f threads/inst" % (len(rows) - 2, ns, ninst, nthr / max(ninst, 1)))
    print("== stall mix: " + ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(ns, 1)) for h, v in tot.most_common(8)))
    print("== opcode mix (inst%, samples%): " + ", ".join("%s %.1f/%.1f" % (op, 100.0 * a / max(ninst, 1), 100.0 * b / max(ns, 1))
                                                         for op, (a, b) in sorted(opc.items(), key=lambda kv: -kv[1][0])[:14]))
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    cur, hdr, out, seen = None, None, [], set()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            if cur is not None and r[1].split("/")[-1] in seen:
                break                   # second launch starts: its files repeat
            cur = r[1].split("/")[-1]; seen.add(cur); continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr and r[0] != "":
            try:
                s, ie, te = int(r[6] or 0), int(r[7] or 0), int(r[8] or 0)
            except ValueError:
                continue
            out.append((s, ie, te, cur, r[0], r[1].strip()[:88]))
    tot_s = sum(o[0] for o in out) or 1
    tot_i = sum(o[1] for o in out) or 1
    print("== hottest source lines (samples%, inst%, threads/inst)")
    for o in sorted(out, key=lambda o: -o[0])[:n_lines]:
        print("  %5.2f%% %5.2f%% %4.1f  %s:%s  %s" % (100.0 * o[0] / tot_s, 100.0 * o[1] / tot_i, o[2] / max(o[1], 1), o[3], o[4], o[5]))


if __name__ == "__main__":
    main()
