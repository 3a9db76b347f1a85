#!/usr/bin/env python
"""SASS evidence for the shipped library (run here, no GPU): per-kernel counts of the tcgen05 / TMA / TMEM mnemonics and an
excerpt of tc_run_kernel<3> -- one MMA-issue sequence of a hidden layer and one 8-column epilogue chunk.
    python tools/sass_excerpt.py > profiles/r02_tc_run_kernel.sass"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mcmc-mems_b200", "libmems_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "MUFU.EX2", "MUFU.RCP", "F2FP", "FHADD", "FFMA2", "FMUL2", "FADD2",
        "LDL", "STL", "BAR.SYNC"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs, name = {}, None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln):
            funcs[name].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", re.sub(r"^\s+/\*[0-9a-f]{4,6}\*/\s+", "", ln)))
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): mnemonic counts per kernel")
    for fn, body in funcs.items():
        short = re.sub(r"\(anonymous namespace\)::", "", fn).split("(")[0]
        cnt = {k: sum(1 for x in body if re.search(r"(^|\s)" + re.escape(k), x)) for k in KEYS}
        if cnt["UTCHMMA"] or cnt["LDTM"] or "tc_" in short:
            print(f"{short:40s} {len(body):6d} instr  " + " ".join(f"{k}={v}" for k, v in cnt.items() if v))
    run3 = next((b for f, b in funcs.items() if "tc_run_kernel<3>" in f), None)
    if run3 is None:
        sys.exit("tc_run_kernel<3> not found")
    i0 = next(i for i, x in enumerate(run3) if "UTCHMMA" in x)
    # the longest run of UTCHMMA-dense code = one hidden layer's 12 + 1 MMAs and the commit
    best, start = 0, i0
    for i, x in enumerate(run3):
        if "UTCHMMA" in x:
            n = sum(1 for y in run3[i:i + 60] if "UTCHMMA" in y)
            if n > best:
                best, start = n, i
    print("\n# tc_run_kernel<3>: MMA issue of one hidden layer (a_lo*W_hi, a_hi*W_lo, a_hi*W_hi: 12 x K=16, + bias) and the commit")
    for x in run3[max(start - 4, 0):start + 64]:
        if any(k in x for k in ("UTCHMMA", "UTCBAR", "ELECT", "SYNCS", "UMOV", "R2UR")):
            print("   ", x)
    j = next(i for i, x in enumerate(run3) if x.startswith("LDTM.x8"))
    print("\n# tc_run_kernel<3>: one 8-column epilogue chunk of a hidden thread (tcgen05.ld -> 2^-|y| -> shared reciprocal -> tanh -> "
          "fp16 hi/lo split -> tcgen05.st)")
    k = next(i for i in range(j + 1, len(run3)) if run3[i].startswith("LDTM"))
    for x in run3[j - 12:k + 1]:
        print("   ", x)
    w = next(i for i, x in enumerate(run3) if "UBLKCP" in x)
    print("\n# tc_run_kernel<3>: weight image staged once per CTA with bulk TMA copies (UBLKCP) on an mbarrier")
    for x in run3[w - 3:w + 4]:
        print("   ", x)


if __name__ == "__main__":
    main()
