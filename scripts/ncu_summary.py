"""Summarise an `ncu --page raw --csv` export: the handful of counters the roofline argument needs."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__sass_thread_inst_executed_op_dfma_pred_on.sum',
        'sm__sass_thread_inst_executed_op_dadd_pred_on.sum', 'sm__sass_thread_inst_executed_op_dmul_pred_on.sum']


def main(path, extra=()):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for k in list(KEYS) + list(extra):
        for i, h in enumerate(hdr):
            if h == k:
                print("%-70s %-12s %s" % (k, units[i], " | ".join(r[i] for r in rows[2:])))
    for i, h in enumerate(hdr):
        if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
            vals = [r[i] for r in rows[2:]]
            try:
                if max(float(v.replace(',', '')) for v in vals) >= 2.0:
                    print("%-70s %-12s %s" % (h.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('smsp__average_warp_latency_issue_stalled_','stall:'), units[i], " | ".join(vals)))
            except ValueError:
                pass
    print("kernels:", [r[hdr.index('Kernel Name')][:60] for r in rows[2:]])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
