#!/usr/bin/env python3
"""Static SASS mnemonic histogram of libcbfrrt.so per kernel (cuobjdump -sass) -> profiles/<tag>_sass_mnemonic_histogram.txt.
   python tools/sass_histogram.py profiles/r2_sass_mnemonic_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cbf-inc-rrt_b200", "cbfrrt", "_lib", "libcbfrrt.so")
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "REDUX", "SHFL", "VOTE", "UCGABAR", "FFMA", "FADD", "FMUL", "FMNMX",
        "FMNMX3", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "MUFU", "DFMA", "ATOMG", "REDG"]


def main():
    out = sys.argv[1]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    kernels, cur, i = [], None, 0
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = [re.sub(r"\(.*", "", names[i]).replace("cbfrrt::", ""), collections.Counter(), 0]
            kernels.append(cur)
            i += 1
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[1][m.group(1)] += 1
            cur[2] += 1
    kernels.sort(key=lambda k: -k[2])
    with open(out, "w") as f:
        f.write("SASS mnemonic histogram of libcbfrrt.so (sm_100a), `cuobjdump -sass`, static instruction counts per kernel "
                "(tools/sass_histogram.py).\n"
                "tcgen05.mma -> UTCHMMA, tcgen05.commit -> UTCBAR, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS, "
                "redux.sync -> REDUX, shfl -> SHFL, ballot/all -> VOTE, barrier.cluster -> UCGABAR;\n"
                "multimem.st lowers to STG.E.STRONG.SYS (no distinct mnemonic).\n\n")
        f.write("kernel | instructions | " + " | ".join(COLS) + "\n")
        for name, c, n in kernels:
            f.write(f"{name} | {n} | " + " | ".join(str(c.get(k, 0)) for k in COLS) + "\n")
    print(len(kernels), "kernels ->", out)


if __name__ == "__main__":
    main()
