import csv,sys,subprocess
rep=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 40
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__t_sector_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum ','dram__bytes_read.sum ','dram__bytes_write.sum ','lts__t_bytes.sum ','l1tex__t_bytes.sum ']
for h,u,v in zip(hdr,units,vals):
    if any(w in h+' ' for w in want): print(h,u,v)
for h,u,v in zip(hdr,units,vals):
    if 'smsp__average_warps_issue_stalled' in h and float(v)>0.2: print(h.replace('smsp__average_warps_issue_stalled_',''),v)
src=subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
cur=None; agg=[]
for r in rows:
    if len(r)==2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)<12 or r[0] in ('Line No',''): continue
    try: ln=int(r[0]); inst=int(r[7]); thr=int(r[8]); smp=int(r[6])
    except: continue
    if inst>0: agg.append((inst,thr,smp,cur,ln,r[1].strip()[:100]))
tot=sum(a[0] for a in agg); tthr=sum(a[1] for a in agg); ts=sum(a[2] for a in agg)
print("total warp inst",tot,"avg thr",tthr/tot)
for a in sorted(agg,key=lambda x:-x[2])[:top]:
    print(f"{a[2]/ts*100:5.1f}% smp {a[0]/tot*100:5.1f}% inst  avgthr {a[1]/a[0]:5.1f}  {a[3]}:{a[4]}  {a[5]}")
