"""Condense `ncu -i rep --page raw --csv` output (tools/gpu.sh ncu) into the per-kernel summary
kept under profiles/: duration, DRAM bytes, tensor / issue activity, L2 RED and read sectors.

    python tools/ncu_summary.py gpurun_out/ncu_TAG_raw.csv > profiles/ncu_TAG_summary.csv
"""

import csv
import sys

WANT = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_red.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
]


def main(path):
    rows = list(csv.reader(line for line in open(path) if line.startswith('"')))
    head, units, data = rows[0], rows[1], rows[2:]
    cols = {}
    for w in WANT:
        for i, h in enumerate(head):
            if h == w or h.endswith("." + w):
                cols[w] = i
                break
    ki = head.index("Kernel Name")
    out = csv.writer(sys.stdout)
    out.writerow(["kernel"] + [f"{w} [{units[cols[w]]}]" for w in WANT if w in cols])
    for r in data:
        out.writerow([r[ki][:90]] + [r[cols[w]] for w in WANT if w in cols])


if __name__ == "__main__":
    main(sys.argv[1])
