import csv,sys,subprocess
rep=sys.argv[1]; nel=float(sys.argv[2])
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; hdr=None; agg={}; ti=ts=0
for r in rows:
    if len(r)==2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>5 and r[0]=='Line No': hdr=r; continue
    if hdr and len(r)==len(hdr) and r[2]=='-':
        try: ln=int(r[0])
        except: continue
        inst=int(r[hdr.index('Instructions Executed')] or 0); samp=int(r[hdr.index('# Samples')] or 0)
        agg[(cur,ln)]=(inst,samp,r[1].strip()[:95]); ti+=inst; ts+=samp
print('warp inst', ti, 'thread-inst/elem %.0f'%(ti*32/nel))
for (f,ln),(inst,samp,src) in sorted(agg.items(), key=lambda kv:-(kv[1][0]/ti+kv[1][1]/ts))[:int(sys.argv[3])]:
    print('%-10s %4d inst %5.2f%% (%.1f/el) samp %5.2f%% %s'%(f[:10],ln,100*inst/ti, inst*32/nel, 100*samp/ts, src))
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr=rows[0]; r=rows[2]
for w in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','lts__t_sectors_op_atom.sum','lts__t_sectors_op_red.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']:
    if w in hdr: print(w, r[hdr.index(w)])
def f(x):
    try: return float(x)
    except: return 0
items=[(h,r[i]) for i,h in enumerate(hdr) if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
print([ (h.replace('smsp__pcsamp_warps_issue_stalled_',''),v) for h,v in sorted(items,key=lambda kv:-f(kv[1]))[:8]])
