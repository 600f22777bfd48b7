"""Print the metrics we track from an .ncu-rep (run where ncu is installed, no GPU needed):
    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep
"""
import csv
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
    'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
    'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_alu.sum',
    'smsp__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_xu.sum',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
    'l1tex__throughput.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
    'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
    'smsp__warps_eligible.avg.per_cycle_active',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        print(f'== {name[:100]}')
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'{w:86s} {vals[i]:>18s} {units[i]}')


if __name__ == '__main__':
    main(sys.argv[1])
