#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` report into the text kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] > profiles/rN_ncu_<what>.txt

Reads the report with `ncu -i ... --page raw --csv` (metrics) and `--page source --csv` (per-instruction
warp-stall samples), and prints, per kernel launch: duration, DRAM bytes, pipe / issue utilisation, occupancy
limits, shared-memory conflicts, the stall-reason mix and the opcodes that hold the most samples.
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True)
    if out.returncode:
        raise SystemExit(out.stderr)
    return out.stdout


def raw_rows(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    head, units = rows[0], rows[1]
    return head, units, rows[2:]


def source_tables(rep):
    """yield (kernel name, list of dict rows) for each kernel of the source page."""
    text = ncu(rep, "source")
    name, buf = None, []
    for line in text.splitlines():
        if line.startswith('"Kernel Name"') or line.startswith("Kernel Name"):
            if buf:
                yield name, list(csv.DictReader(io.StringIO("\n".join(buf))))
            name, buf = line.split(",", 1)[-1].strip('"'), []
        elif line.startswith('"') and name is not None:
            buf.append(line)
    if buf:
        yield name, list(csv.DictReader(io.StringIO("\n".join(buf))))


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    head, units, rows = raw_rows(rep)
    col = {h: i for i, h in enumerate(head)}
    for r in rows:
        kname = r[col["Kernel Name"]]
        if want not in kname:
            continue
        print(kname)
        for m in METRICS:
            if m in col:
                print("  %-80s %s %s" % (m, r[col[m]], units[col[m]]))
    try:
        for kname, tab in source_tables(rep):
            if want not in (kname or "") or not tab:
                continue
            keys = tab[0].keys()
            samp_key = next((k for k in keys if k.startswith("# Samples") or k == "Warp Stall Sampling (All Samples)"), None)
            src_key = next((k for k in keys if k in ("Source", "SASS")), None)
            stall_keys = [k for k in keys if k.startswith("stall_")]
            if not samp_key or not src_key:
                continue
            tot = collections.Counter(); ops = collections.Counter(); nsamp = 0
            for row in tab:
                try:
                    n = int(float(row[samp_key] or 0))
                except ValueError:
                    continue
                nsamp += n
                op = (row[src_key].split() or ["?"])[0]
                if op.startswith("@"):
                    op = (row[src_key].split() + ["?"])[1]
                ops[op.split(".")[0]] += n
                for k in stall_keys:
                    try:
                        tot[k] += int(float(row[k] or 0))
                    except ValueError:
                        pass
            print("samples", nsamp)
            if tot:
                s = sum(tot.values()) or 1
                print(" stalls: " + ", ".join("%s %d%%" % (k[6:], 100 * v // s) for k, v in tot.most_common(8)))
            print(" ops by samples: " + ", ".join("%s %d%%" % (k, 100 * v // max(nsamp, 1)) for k, v in ops.most_common(14)))
    except Exception as e:                      # the source page layout differs between ncu versions; metrics still print
        print("# source page not summarised: %r" % (e,))


if __name__ == "__main__":
    main()
