#!/usr/bin/env python
"""Text summary of one `ncu --set full` capture (profiles/ keeps these, gpurun_out/ keeps the .ncu-rep):
duration, DRAM / L2 traffic and throughput, tensor / FMA / XU pipe utilisation, issue rate, occupancy, and the
warp-state (stall) sample histogram.

    python scripts/ncu_summary.py gpurun_out/f_sinkhorn_v2.ncu-rep > profiles/r02_prof_sinkhorn_v2_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_inst_executed_op_tmem_ldt.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        print(f"== {d.get('Kernel Name', ('', '?'))[1]}   grid {d.get('Grid Size', ('', '?'))[1]} block {d.get('Block Size', ('', '?'))[1]}")
        for k in KEYS:
            if k in d and d[k][1] != "":
                print(f"   {k:<72} {d[k][1]:>18} {d[k][0]}")
        st = [(h[len('smsp__pcsamp_warps_issue_stalled_'):], float(v.replace(',', ''))) for h, (u, v) in d.items()
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and v not in ("", "0")]
        tot = sum(v for _, v in st) or 1.0
        print("   warp-state samples: " + ", ".join(f"{n} {100 * v / tot:.1f}%" for n, v in sorted(st, key=lambda x: -x[1])[:8]))


if __name__ == "__main__":
    main()
