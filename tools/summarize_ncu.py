#!/usr/bin/env python3
"""Turn one ncu report (gpurun_out/*.ncu-rep) into the text summaries kept under profiles/:
   <out>_details.txt (ncu --page details), <out>_raw.csv (--page raw), <out>_lines.txt (executed warp instructions and
   stall samples per source file and per source line, from --page source).
   python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1c_ncu_full_safety_tc"""
import collections, csv, io, subprocess, sys


def run(*a):
    return subprocess.run(["ncu", "-i", sys.argv[1], *a], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    open(out + "_details.txt", "w").write(run("--page", "details"))
    open(out + "_raw.csv", "w").write(run("--page", "raw", "--csv"))
    rows = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "cuda,sass"))))
    I = lambda x: int(x) if x.lstrip("-").isdigit() else 0
    sec = hdr = None
    files, lines = collections.Counter(), collections.defaultdict(lambda: [0, 0])
    fsamp = collections.Counter()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            sec, hdr = r[1], None
        elif r[0] == "Line No":
            hdr, ie, isamp = r, r.index("Instructions Executed"), r.index("# Samples")
        elif hdr is not None and r[0].isdigit():
            files[sec] += I(r[ie]); fsamp[sec] += I(r[isamp])
            lines[(sec, int(r[0]))][0] += I(r[ie]); lines[(sec, int(r[0]))][1] += I(r[isamp])
    tot, ts = max(1, sum(files.values())), max(1, sum(fsamp.values()))
    with open(out + "_lines.txt", "w") as f:
        f.write("executed warp instructions / stall samples per source file\n")
        for k, v in files.most_common(8):
            f.write(f"{v:12d} {100 * v / tot:5.1f}%  samples {100 * fsamp[k] / ts:5.1f}%  {k}\n")
        f.write("\ntop source lines\n")
        for (k, ln), (n, sm) in sorted(lines.items(), key=lambda x: -x[1][0])[:60]:
            try:
                txt = open(k).read().split("\n")[ln - 1].strip()[:110]
            except Exception:
                txt = ""
            f.write(f"{k.split('/')[-1]:14s} {ln:5d} instr {100 * n / tot:5.2f}% samples {100 * sm / ts:5.2f}%  {txt}\n")


if __name__ == "__main__":
    main()
