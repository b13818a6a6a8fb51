#!/usr/bin/env python
"""SASS evidence for the instruction-level claims of DESIGN.md (runs here: cuobjdump needs no GPU).
usage: python scripts/sass_excerpts.py > profiles/r02_sass_excerpts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tropical_cyclone_risk_b200", "libtcrisk.so")
CLAIMS = (
    ("k_fourier_table_mma", ("DMMA", "LDGSTS"), "FP64 tensor-core tabulation, cp.async coefficient tiles"),
    ("k_integrateILi192ELi2ELi2ELi319", ("LDG.E.ENL2.256", "MUFU.RSQ64H", "MUFU.RCP64H", "BAR.", "DFMA", "F2F.F64.F32", "CALL"),
     "default integrator (variant 22): 256-bit gathers, nvcc's own sqrt / rcp seeds inline, slot barriers"),
    ("k_env_interpILi3", ("LDG.E.128", "STG.E"), "stand-alone sampler: 128-bit record reads"),
    ("k_env_interp_asyncILi256", ("LDGSTS",), "sampler variant 5: cp.async burst per tile"),
    ("k_wind_stats_singleILi64ELi128", ("LDGSTS.E.BYPASS.128", "LDGSTS"), "wind statistics: 16-byte cp.async of the month's samples"),
    ("15k_fourier_table6", ("LDGSTS", "DFMA"), "scalar Fourier tabulation (A/B variant)"),
)


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
            funcs[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip())
    print("# cuobjdump -sass tropical_cyclone_risk_b200/libtcrisk.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -fmad=false)")
    for key, ops, what in CLAIMS:
        name = next((f for f in funcs if key in f), None)
        if not name:
            print("\n## %s: NOT FOUND" % key)
            continue
        body = funcs[name]
        print("\n## %s -- %s\n# %d SASS instructions" % (name, what, len(body)))
        for op in ops:
            hits = [l for l in body if op in l]
            print("# %-22s %5d occurrences" % (op, len(hits)))
            for l in hits[:3]:
                print("   " + l.strip())
    blocks = []
    name = next(f for f in funcs if "k_integrateILi192ELi2ELi2ELi319" in f)
    n = 0
    for l in funcs[name]:
        if re.search(r"\b(BRA|BSSY|BSYNC|CALL|EXIT|RET|BAR)\b", l):
            blocks.append(n); n = 0
        else:
            n += 1
    print("\n## basic blocks of the default integrator: longest %d instructions (the straight-line RHS), next %s" % (
        max(blocks), sorted(blocks, reverse=True)[1:5]))


if __name__ == "__main__":
    main()
