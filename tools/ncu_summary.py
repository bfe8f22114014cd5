#!/usr/bin/env python3
"""tools/ncu_summary.py <report.ncu-rep> <out.md> [title] -- key metrics of an `ncu --set full` capture
(read on the CPU box with `ncu -i ... --page raw --csv`) as a small markdown table for profiles/."""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
    "dram__cycles_elapsed.avg.per_second",
]
with open(out, "w") as fh:
    fh.write("# %s\n\n" % title)
    fh.write("Source: `ncu --set full --clock-control none --import-source on` on a B200, read with "
             "`ncu -i <rep> --page raw --csv` (tools/ncu_summary.py).  One column per captured launch.\n\n")
    fh.write("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |\n")
    fh.write("|---|---|" + "---|" * len(data) + "\n")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            fh.write("| %s | %s | %s |\n" % (w, units[i], " | ".join(r[i] for r in data)))
    try:
        ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        fh.write("\n")
        for n, r in enumerate(data):
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tb = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
            tus = float(r[it]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}[units[it]]
            fh.write("- launch %d: DRAM traffic %.4f GB in %.1f us = %.0f GB/s\n" % (n, tb / 1e9, tus, tb / 1e9 / (tus * 1e-6)))
    except Exception as e:
        fh.write("\n(traffic summary failed: %r)\n" % (e,))
print(open(out).read())
