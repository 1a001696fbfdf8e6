"""Prints the metrics we track from an .ncu-rep (run here, no GPU needed): python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','launch__grid_size','launch__occupancy_limit_registers','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors_srcunit_tex_op_read.sum','smsp__warps_eligible.avg.per_cycle_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_xu.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed_op_branch.sum']
for vals in rows[2:]:
    print('-----')
    for i, h in enumerate(hdr):
        if h in want:
            print(f"{h} [{units[i]}] = {vals[i]}")
        elif 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct'):
            try:
                if float(vals[i]) >= 4: print(f"{h} = {vals[i]}")
            except ValueError: pass
