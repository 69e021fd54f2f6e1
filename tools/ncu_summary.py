#!/usr/bin/env python
"""One line per kernel launch of an `ncu -i x.ncu-rep --page raw --csv` dump: the handful of metrics the roofline
discussion in DESIGN.md uses (duration, block / grid size, tensor-pipe activity, DRAM bytes, L2 / shared-memory throughput).

    python tools/ncu_summary.py gpurun_out/r2_13_body_raw.csv > profiles/ncu_body_r2.txt
"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "us"),
    ("launch__block_size", "blk"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%elapsed"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma%act"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("dram__bytes_read.sum", "dramR"),
    ("dram__bytes_write.sum", "dramW"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
    ("sm__cycles_active.avg", "sm_cycles"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {c: i for i, c in enumerate(hdr)}
    ki = idx["Kernel Name"]
    for r in rows[2:]:
        name = r[ki].replace("void ", "").replace("<unnamed>::", "").split("(")[0][:44]
        parts = [f"{name:44s}"]
        for col, short in COLS:
            if col in idx:
                v, u = r[idx[col]], units[idx[col]]
                parts.append(f"{short}={v}{u if u in ('us', 'Mbyte', 'Kbyte', 'byte', 'Gbyte') else ''}")
        print("  ".join(parts))


if __name__ == "__main__":
    main(sys.argv[1])
