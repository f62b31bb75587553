"""Summarise an `ncu --page raw --csv` dump: python tools/ncu_summary.py raw.csv"""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_static',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__cycles_elapsed.avg',
        'sm__cycles_elapsed.avg.per_second', 'smsp__cycles_active.avg',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_drain_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct']


def main(path, grep=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('----')
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('%-72s %s %s' % (w, r[i][:90], units[i]))
        if grep:
            for i, h in enumerate(hdr):
                if grep in h and h not in WANT:
                    print('%-72s %s %s' % (h, r[i], units[i]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
