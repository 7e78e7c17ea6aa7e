#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read on the CPU box): one column per captured launch, the metrics the
roofline discussion needs.   python tools/ncu_summary.py gpurun_out/x.ncu-rep [--md]"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu (SFU) pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of max"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    names = [r[ki].split("(")[0].replace("void ngpde::", "")[:34] for r in data]
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in WANT:
        if key not in hdr:
            continue
        i = hdr.index(key)
        print(f"| {label} ({units[i]}) | " + " | ".join(r[i] for r in data) + " |")


if __name__ == "__main__":
    main()
