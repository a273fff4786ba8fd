import csv,sys,subprocess
rep=sys.argv[1]; kern=sys.argv[2] if len(sys.argv)>2 else None
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
idx={h:i for i,h in enumerate(hdr)}
W=['gpu__time_duration.sum','inst_executed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','lts__t_sectors_srcunit_tex_op_red.sum','dram__bytes_read.sum','dram__bytes_write.sum','sass__inst_executed_local_loads','lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for r in rows[2:]:
    if kern and kern not in r[idx['Kernel Name']]: continue
    print('==',r[idx['Kernel Name']][:80])
    for w in W:
        if w in idx: print('  ',w,'=',r[idx[w]],units[idx[w]])
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv']+(['--kernel-name','regex:'+kern] if kern else []),capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=[r for r in rows[2:] if len(r)==len(hdr)]
iE=hdr.index('Instructions Executed'); iS=hdr.index('Source'); iT=hdr.index('Avg. Threads Executed'); iSm=hdr.index('# Samples')
tot=sum(int(r[iE]) for r in data)
print('total warp inst', tot, 'n instr', len(data))
seg=int(sys.argv[3]) if len(sys.argv)>3 else 200
for s in range(0,len(data),seg):
    chunk=data[s:s+seg]
    e=sum(int(r[iE]) for r in chunk); sm=sum(int(r[iSm]) for r in chunk)
    at=sum(float(r[iT])*int(r[iE]) for r in chunk)/max(e,1)
    if e: print(f'{s:5d}-{s+seg:5d}: inst {e/tot*100:5.1f}%  samples {sm:6d} thr {at:5.1f}  first: {chunk[0][iS].strip()[:36]:36s} maxexec {max(int(r[iE]) for r in chunk)}')
