#!/usr/bin/env python
"""One line per captured launch of an .ncu-rep (read here, no GPU needed): time, DRAM bytes, achieved GB/s, occupancy, hit rates.
usage: python scripts/ncu_table.py <report.ncu-rep> [out.txt]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]

    def col(r, k, scale=None):
        i = hdr.index(k)
        v = float(r[i].replace(",", ""))
        u = units[i]
        if scale == "us":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(u, 1.0)
        if scale == "MB":
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
        return v

    print("# GB/s = measured DRAM bytes / kernel time (cold-cache, serialised ncu --set full capture)", file=out)
    print("%-38s %7s %8s %8s %8s %6s %6s %5s %6s %6s" % ("kernel", "us", "rdMB", "wrMB", "GB/s", "dram%", "regs", "occ%", "l1hit", "l2hit"), file=out)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")][:38]
        us = col(r, "gpu__time_duration.sum", "us")
        rd, wr = col(r, "dram__bytes_read.sum", "MB"), col(r, "dram__bytes_write.sum", "MB")
        print("%-38s %7.1f %8.1f %8.1f %8.0f %6.1f %6d %5.1f %6.1f %6.1f" % (
            name, us, rd, wr, (rd + wr) * 1e6 / (us * 1e-6) / 1e9,
            col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), int(col(r, "launch__registers_per_thread")),
            col(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), col(r, "l1tex__t_sector_hit_rate.pct"),
            col(r, "lts__t_sector_hit_rate.pct")), file=out)


if __name__ == "__main__":
    main()
