#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters the DESIGN/roofline discussion uses."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__waves_per_multiprocessor','launch__grid_size',
 'launch__block_size','launch__shared_mem_per_block_dynamic','sm__cycles_elapsed.max',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
 'smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'smsp__warps_eligible.avg.per_cycle_active','sm__maximum_warps_per_active_cycle_pct',
 'lts__t_sectors.sum','lts__t_bytes.sum','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum','l1tex__data_pipe_lsu_wavefronts.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum','smsp__inst_executed_pipe_lsu.sum',
 'lts__t_sectors_srcunit_tex.sum','lts__t_sectors_srcunit_tex_lookup_hit.sum','lts__t_sectors_srcunit_tex_lookup_miss.sum']
def main(rep):
    out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
        print('kernel:', name)
        for i,h in enumerate(hdr):
            if h in WANT: print(f"  {h:88s} {units[i]:16s} {r[i]}")
if __name__ == '__main__':
    main(sys.argv[1])
