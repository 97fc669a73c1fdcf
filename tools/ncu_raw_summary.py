#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export into one line per kernel launch (the metrics DESIGN.md / bench.py quote):
duration, DRAM bytes read + written, DRAM throughput %, SM throughput %, issue-slot utilisation, achieved occupancy,
registers, dynamic + static shared memory, tensor-pipe activity.    python tools/ncu_raw_summary.py in.csv > out.csv"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "duration_ns"), ("dram__bytes_read.sum", "dram_read_B"), ("dram__bytes_write.sum", "dram_write_B"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue_pct"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem_dyn_B"), ("launch__shared_mem_per_block_static", "smem_static_B"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"), ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("smsp__inst_executed.sum", "warp_inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes_per_inst"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("lts__t_bytes.sum", "l2_bytes")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    unit_scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "ns": 1, "us": 1e3, "ms": 1e6, "usecond": 1e3, "msecond": 1e6, "nsecond": 1, "second": 1e9}
    out = csv.writer(sys.stdout)
    out.writerow(["kernel"] + [w[1] for w in WANT])
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        line = [r[col["Kernel Name"]].split("(")[0]]
        for key, _ in WANT:
            if key not in col:
                line.append("")
                continue
            v = r[col[key]].replace(",", "")
            try:
                x = float(v) * unit_scale.get(units[col[key]], 1)
                line.append("%.6g" % x)
            except ValueError:
                line.append(v)
        out.writerow(line)


if __name__ == "__main__":
    main(sys.argv[1])
