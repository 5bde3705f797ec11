#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i`) into the handful of numbers the design discussion uses.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print("kernel:", d.get("Kernel Name", ("", "?"))[1])
        for k in KEYS:
            if k in d:
                print(f"  {k:95s} {d[k][1]:>18s} {d[k][0]}")
        tensor = {k: v for k, v in d.items() if "tensor" in k and "pct_of_peak_sustained_active" in k and ".avg." in k}
        for k, v in sorted(tensor.items()):
            if k not in KEYS:
                print(f"  {k:95s} {v[1]:>18s} {v[0]}")
        stalls = {k: float(v[1]) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v[1]}
        print("  warp stall reasons (warps stalled per issue slot):")
        for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]:
            print(f"    {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:8.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
