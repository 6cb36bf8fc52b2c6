#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): per captured launch the counters the roofline needs.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % (inst)"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe % (cycles)"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__occupancy_limit_registers", "occ limit regs (blocks)"),
        ("launch__occupancy_limit_shared_mem", "occ limit smem (blocks)"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"), ("sm__cycles_elapsed.avg", "SM cycles")]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    print(f"# {rep}: ncu --set full --clock-control none, one column per captured launch")
    for r in rows[2:]:
        print("kernel:", r[name_i][:110])
    for k, label in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{label:28s} [{units[i]:>14s}] " + "  ".join(f"{r[i]:>14s}" for r in rows[2:]))
    ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")

    def tobytes(v, u):
        m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        return float(v.replace(",", "")) * m
    print("traffic (read+write) bytes    " + "  ".join(f"{tobytes(r[ri], units[ri]) + tobytes(r[wi], units[wi]):14.0f}" for r in rows[2:]))


if __name__ == "__main__":
    main()
