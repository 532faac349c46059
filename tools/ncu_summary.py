#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    tensor_like = [h for h in hdr if "tensor" in h and h not in KEYS]
    print(f"# ncu summary of {rep} (ncu --set full --clock-control none; cold-cache, serialised replays)")
    for n, r in enumerate(rows[2:]):
        print(f"\n## launch {n}: {r[hdr.index('Kernel Name')]}")
        for k in KEYS + tensor_like[:6]:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:75s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main()
