#!/usr/bin/env python
"""Hot CUDA source lines of a kernel in an `ncu --set full --import-source on` report (built with -lineinfo):

    python tools/ncu_hotlines.py report.ncu-rep kernel-regex [top N]

Aggregates the per-instruction warp-stall samples and executed-instruction counts of `--page source --print-source
cuda,sass` by source line."""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
    kern = fpath = hdr = None
    agg = collections.defaultdict(list)
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            kern = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0] and len(r) >= 8:
            try:
                agg[kern].append((fpath, int(r[0]), r[1], int(r[6]), int(r[7])))
            except ValueError:
                pass
    for k, v in agg.items():
        ts, ti = sum(x[3] for x in v) or 1, sum(x[4] for x in v) or 1
        print("%s\n  samples %d  warp instructions %d" % (k, ts, ti))
        for f, l, src, s, i in sorted(v, key=lambda x: -x[3])[:top]:
            print("  %-18s %4d  smp %5.1f%%  instr %5.1f%%  %s" % (f, l, 100 * s / ts, 100 * i / ti, src.strip()[:120]))


if __name__ == "__main__":
    main()
