import csv,sys,subprocess
out=subprocess.run(["ncu","-i",sys.argv[1],"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
h,u,v=rows[0],rows[1],rows[2]
d={a:(b,c) for a,b,c in zip(h,u,v)}
for a in sorted(d):
    if ('average_warps_issue_stalled' in a and 'not_issued' not in a and a.endswith('.ratio')) or a in ('gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.per_cycle_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'):
        print(a.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio',''), d[a])
