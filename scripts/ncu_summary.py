"""Print the key metrics of an .ncu-rep (first kernel) -- used to write profiles/*.md summaries."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum','sm__cycles_elapsed.avg','sm__cycles_elapsed.avg.per_second','dram__bytes_read.sum','dram__bytes_write.sum',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
 'launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__shared_mem_config_size','launch__shared_mem_per_block_dynamic',
 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','launch__grid_size','launch__block_size','launch__waves_per_multiprocessor',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed',
 'smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_global_red.sum','lts__t_sectors_op_red.sum','lts__t_sectors_op_atom.sum',
 'lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum']
def main(path, extra=()):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    for row in r[2:]:
        print('## kernel:', row[hdr.index('Kernel Name')][:80])
        for i,h in enumerate(hdr):
            if h in WANT or h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') or any(e in h for e in extra):
                print(f'{h:95s} {row[i]:>20s} {units[i]}')
if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2:])
