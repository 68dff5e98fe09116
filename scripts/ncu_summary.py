"""Summarise an .ncu-rep (ncu --set full) as a markdown table per launch: python scripts/ncu_summary.py report.ncu-rep [title]"""
import csv, subprocess, sys, io
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_barrier',
        'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_lg_throttle', 'smsp__pcsamp_warps_issue_stalled_mio_throttle',
        'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_branch_resolving', 'smsp__pcsamp_warps_issue_stalled_membar',
        'smsp__pcsamp_warps_issue_stalled_not_selected', 'smsp__pcsamp_sample_count']
def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    if len(sys.argv) > 2:
        print('# ' + sys.argv[2] + '\n')
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')].split('(')[0].replace('<unnamed>::', '').replace('void ', '')
        print(f'## {name}\n\n| metric | value | unit |\n|---|---|---|')
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'| {w} | {r[i]} | {units[i]} |')
        print()
main()
