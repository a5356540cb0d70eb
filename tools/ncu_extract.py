#!/usr/bin/env python
"""Prints the judged metrics of one or more .ncu-rep captures as markdown tables (profiles/rN_ncu_summary.md is assembled from
this): python tools/ncu_extract.py gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "gpc__cycles_elapsed.avg.per_second",
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
            print("\n### %s -- `%s`\n" % (path.split("/")[-1], d.get("Kernel Name", ("", "?"))[1]))
            print("| metric | value |\n|---|---|")
            for k in WANT:
                if k in d:
                    print("| %s | %s %s |" % (k, d[k][1], d[k][0]))
            try:
                mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tr = sum(float(d[k][1]) * mult[d[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                print("| traffic = dram read + write | %.1f Mbyte |" % (tr / 1e6))
            except (KeyError, ValueError):
                pass


if __name__ == "__main__":
    main()
