#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into the handful of metrics the
roofline argument needs.  usage: python profiles/summarize_ncu.py gpurun_out/X.ncu-rep > profiles/X.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("kernel:", r[name_col])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-78s %s %s" % (k, r[i], units[i]))
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(r[hdr.index("dram__bytes_write.sum")]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_write.sum")]]
            t = float(r[hdr.index("gpu__time_duration.sum")]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}[units[hdr.index("gpu__time_duration.sum")]]
            print("  traffic (dram read+write) = %.1f MB per launch; %.0f GB/s under ncu (cold cache, serialised)" % ((rd + wr) / 1e6, (rd + wr) / t / 1e9))
        except Exception as e:  # noqa
            print("  (no traffic summary: %s)" % e)


if __name__ == "__main__":
    main(sys.argv[1])
