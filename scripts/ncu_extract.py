#!/usr/bin/env python
"""Extract the per-launch metrics the docs and bench.py quote from an
`ncu --set full` report:

    python scripts/ncu_extract.py gpurun_out/prof_edge_r02.ncu-rep profiles/r02_b_prof_edge_ncu_full.csv

Output: header row, units row, one row per captured launch (the layout of the
round-1 extracts; bench.py's `roofline.traffic` reads dram__bytes_* from the
newest `profiles/*edge*_ncu_full.csv`)."""
import csv
import subprocess
import sys

COLS = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum',
        'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'lts__t_sectors.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active']


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {c: hdr.index(c) for c in COLS if c in hdr}
    kname = hdr.index('Kernel Name')
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['Kernel Name'] + list(idx))
        w.writerow([''] + [units[i] for i in idx.values()])
        for r in rows[2:]:
            if len(r) == len(hdr):
                w.writerow([r[kname].replace('pvs::', '')] + [r[i] for i in idx.values()])
    print(out, len(rows) - 2, 'launches')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
