import csv,sys,subprocess,re,collections
rep=sys.argv[1]; kname=sys.argv[2]; nsamp=float(sys.argv[3])
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv","-k","regex:"+kname,"-c","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]; hdr0=hdr
want=['gpu__time_duration.sum','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__warps_active.avg.per_cycle_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','smsp__average_warp_latency_per_inst_issued.ratio','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','Block Size','Grid Size','Kernel Name']
out=[]
for h,u,v in zip(hdr,units,vals):
    if h in want or (h.startswith('smsp__average_warps_issue_stalled') and float(v or 0)>0.03): out.append(f"{h} [{u}] = {v}")
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass","-k","regex:"+kname,"-c","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]
ia=hdr.index("Instructions Executed"); isrc=hdr.index("Source")
tot=collections.Counter(); total=0
for r in data:
    try: n=int(r[ia])
    except: continue
    op=re.sub(r'^@!?U?P\d+\s+','',r[isrc].strip()).split()[0].split('.')[0]
    tot[op]+=n; total+=n
raw_total=float(vals[hdr0.index('smsp__inst_executed.sum')])
scale=raw_total/total   # the source page may aggregate several launches of the same kernel
tot=collections.Counter({k:v*scale for k,v in tot.items()}); total=raw_total
out.append(f"thread-instructions per sample = {total*32/nsamp:.1f}")
fp=sum(tot[k] for k in ('DFMA','DMUL','DADD','DSETP'))
out.append(f"FP64-pipe instructions per sample = {fp*32/nsamp:.1f} ({fp/total*100:.1f}% of all)")
out.append("per-sample instruction mix: "+str({k: round(v*32/nsamp,1) for k,v in tot.most_common(24)}))
print("\n".join(out))
