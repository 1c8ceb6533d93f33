"""Trim an `ncu --page raw --csv` export to the metrics the roofline discussion uses.
usage: python tools/ncu_extract.py gpurun_out/prof_fwd_raw.csv profiles/out.csv"""
import csv
import sys

KEEP = ['ID', 'Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__memory_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__cycles_active.avg', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']


def main(src, dst):
    rd = list(csv.reader(open(src)))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows:
            w.writerow([r[i] for i in idx])


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
