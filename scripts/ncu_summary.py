"""Summarise an .ncu-rep (raw page) into the handful of numbers we track per kernel."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem',
 'launch__occupancy_limit_registers','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__cycles_active.avg','sm__cycles_elapsed.max',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','launch__grid_size','launch__block_size',
 'dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum',
 'smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_shared_st.sum','smsp__inst_executed_op_global_ld.sum','smsp__inst_executed_op_global_st.sum',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','sm__sass_thread_inst_executed_op_ffma_pred_on.sum']
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w:75s}", [r[i][:28] for r in rows[2:]], rows[1][i])
print("-- stall reasons (per issue) --")
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and 'ratio' in h and 'warps_issue' in h:
        vals = [r[i] for r in rows[2:]]
        try:
            if max(float(v.replace(',', '')) for v in vals) > 0.15:
                print(f"{h.replace('smsp__average_warps_issue_stalled_','')[:45]:45s}", vals)
        except ValueError:
            pass
