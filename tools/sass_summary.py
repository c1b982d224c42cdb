"""SASS / resource listing of the element kernels for profiles/: per kernel the counts of the instructions that
matter for this path (DMMA = FP64 tensor-core mma.sync m8n8k4, DFMA / DMUL / DADD, LDS / STS, LDGSTS = cp.async, UBLKCP = TMA bulk copy, SYNCS = mbarrier,
RED / ATOM, local-memory LDL / STL = spills, BAR), plus registers / shared memory / stack from the cubin.
    python tools/sass_summary.py > profiles/r02_sass_summary.md        (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "4c_b200", "lib4c_b200_solid.so")
WANT = [("hex8 headline <CoupNeoHooke, STIFF>", r"k_hex8_eval_assembleILi0ELb1ELi0ELb0E"),
        ("hex8 headline, TMA map staging", r"k_hex8_eval_assembleILi0ELb1ELi0ELb1E"),
        ("hex8 F-bar", r"k_hex8_eval_assembleILi0ELb1ELi1ELb0E"),
        ("hex8 MAT_VISCONEOHOOKE", r"k_hex8_eval_assembleILi5ELb1ELi0ELb0E"),
        ("hex8 EAS mild, tensor-core kernel (default)", r"k_hex8_eas_mma_eval_assembleILi0ELb1ELi9E"),
        ("hex8 EAS full, tensor-core kernel (default)", r"k_hex8_eas_mma_eval_assembleILi0ELb1ELi21E"),
        ("hex8 EAS mild, FMA kernel (option eas_mma = 0)", r"k_hex8_eas_eval_assembleILi0ELb1ELi9E"),
        ("hex8 EAS full, FMA kernel (option eas_mma = 0)", r"k_hex8_eas_eval_assembleILi0ELb1ELi21E"),
        ("tet10 <iso/vol, STIFF>", r"k_tet10_eval_assembleILi1ELb1ELi0E"),
        ("hex27 <SVK, STIFF>", r"k_hex27_eval_assembleILi2ELb1ELi0E")]
OPS = ["DMMA", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDGSTS", "UBLKCP", "SYNCS", "LDG", "STG", "RED", "ATOM", "LDL", "STL",
       "BAR", "SHFL", "WARPSYNC"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    funcs = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            funcs[name]["_total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + ".") or (o in ("RED", "ATOM") and op.startswith(o + "G.")):
                    funcs[name][o] += 1
    print("# SASS view of the element kernels (cuobjdump -sass / -res-usage of lib4c_b200_solid.so, sm_100a)\n")
    print("Static instruction counts per kernel (not executed counts); LDL/STL = local memory (spills).\n")
    print("| kernel | resources | total | " + " | ".join(OPS) + " |")
    print("|---|---|---|" + "---|" * len(OPS))
    for label, pat in WANT:
        hits = [n for n in funcs if re.search(pat, n)]
        if not hits:
            print(f"| {label} | (not found: {pat}) |")
            continue
        n = hits[0]
        c = funcs[n]
        print(f"| {label} `{n[:60]}` | {usage.get(n, '?')} | {c['_total']} | " + " | ".join(str(c[o]) for o in OPS) + " |")
    print("\nRED = `REDG.E.ADD.F64.RN.STRONG.GPU` (red.global.add.f64), LDGSTS = cp.async, UBLKCP.S.G + SYNCS.ARRIVE.TRANS64 / "
          "SYNCS.PHASECHK.TRANS64.TRYWAIT = cp.async.bulk + mbarrier (TMA).  FP64 tensor-core DMMA is absent on purpose "
          "(measured: same pipe, no gain, DESIGN 3.1); tcgen05 has no FP64 kind.")
    # a short excerpt of the hot loops as evidence
    for label, pat, grep in (("hex8 headline: phase-2 DFMA chain and the row-wise RED", WANT[0][1], r"DFMA|REDG"),
                             ("hex8 TMA variant: bulk copy + mbarrier", WANT[1][1], r"UBLKCP|SYNCS|LDGSTS")):
        hits = [n for n in funcs if re.search(pat, n)]
        if not hits:
            continue
        body = sass.split("Function : " + hits[0], 1)[1].split("Function : ", 1)[0]
        lines = [l.rstrip() for l in body.splitlines() if re.search(grep, l)]
        print(f"\n## {label}\n\n```")
        for l in lines[:14]:
            print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l)[:140])
        print(f"... ({len(lines)} matching lines)\n```")


if __name__ == "__main__":
    sys.exit(main())
