#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.
usage: python scripts/ncu_summary.py <report.ncu-rep> [out.txt]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__cycles_elapsed.max"]


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== launch:", r[hdr.index("Kernel Name")], file=out)
        for k in KEYS:
            if k in hdr:
                print("  %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]), file=out)
        stalls = [(h, float(r[i])) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
        stalls.sort(key=lambda t: -t[1])
        print("  stall reasons (warps per issue-active cycle):", file=out)
        for h, v in stalls[:8]:
            print("    %-40s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v), file=out)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = next((r for r in rows if "Source" in r and "Instructions Executed" in r), None)
    if hdr:
        iS, iE, iN, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        ops, samp, thr = collections.Counter(), collections.Counter(), collections.Counter()
        for r in rows:
            if len(r) <= iT:
                continue
            t = r[iS].split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).rstrip(";")
            op = ".".join(op.split(".")[:2])
            try:
                e, s, th = int(r[iE]), int(r[iN]), int(r[iT])
            except ValueError:
                continue
            ops[op] += e
            samp[op] += s
            thr[op] += th
        tot, ts = sum(ops.values()) or 1, sum(samp.values()) or 1
        print("== SASS opcode mix (all captured launches): %d instruction lines executed %d warp-instr" % (len(rows), tot), file=out)
        for k, v in ops.most_common(25):
            print("  %-20s %5.1f%% of instr  %5.1f%% of samples  avg active threads %.1f" %
                  (k, 100.0 * v / tot, 100.0 * samp[k] / ts, thr[k] / max(v, 1)), file=out)


if __name__ == "__main__":
    main()
