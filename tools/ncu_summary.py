#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the CSV kept under profiles/.  usage: ncu_summary.py rep out.csv"""
import csv
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum']


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for ri, vals in enumerate(rows[2:]):
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
            f.write("# launch %d: %s\nmetric,unit,value\n" % (ri, name))
            for h, u, v in zip(hdr, units, vals):
                if h in KEEP or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')):
                    f.write("%s,%s,%s\n" % (h, u, v))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
