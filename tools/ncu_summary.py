import subprocess, csv, io, sys
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__grid_size','sm__cycles_elapsed.avg']
idx={h:i for i,h in enumerate(hdr)}
for r in rows[2:]:
    print('----', r[idx['Kernel Name']][:90])
    for w in want[1:]:
        if w in idx: print(f'   {w:75s} {r[idx[w]]} {units[idx[w]]}')
    # stall reasons
    st=[(h,float(r[i].replace(',',''))) for h,i in idx.items() if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and r[i] not in ('','n/a')]
    st.sort(key=lambda t:-t[1])
    print('   top stalls:', [(h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),round(v,2)) for h,v in st[:6]])
