"""Source-level summary of an `ncu --set full --import-source on` report (run where ncu is installed; no GPU needed).

    python profiles/r02/tools/ncu_source_hotspots.py REPORT.ncu-rep [LINES]

Prints the headline metrics, the stall reasons per issued instruction, the share of instructions / stall samples / active
threads per source file, and the LINES source lines with the most stall samples (with their three largest stall reasons).
This is the view that found the serialised carry copy, the parameter-row loads and the idle line-search lanes of round 2.
"""
import csv, sys, subprocess
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); d=dict(zip(rows[0],rows[2]))
for k in ['gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.per_cycle_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__thread_inst_executed_per_inst_executed.ratio','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct']:
    print(k, d.get(k))
for k,v in d.items():
    if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and float(v)>0.05: print('  ',k.split('stalled_')[1].split('_per')[0], v)
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass,cuda'],capture_output=True,text=True).stdout
cur=None; ix=None; agg={}
for r in csv.reader(src.splitlines()):
    if not r: continue
    if r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': ix={h:i for i,h in enumerate(r)}; continue
    if cur is None or ix is None or r[0]=='' : continue
    try: ln=int(r[0]); ie=float(r[ix['Instructions Executed']]); smp=float(r[ix['# Samples']]); te=float(r[ix['Thread Instructions Executed']])
    except: continue
    a=agg.setdefault((cur,ln),[0,0,r[1],{},0]); a[0]+=ie; a[1]+=smp; a[4]+=te
    for k in ix:
        if k.startswith('stall_') and 'Not Issued' not in k:
            try: a[3][k]=a[3].get(k,0)+float(r[ix[k]])
            except: pass
tot=sum(v[0] for v in agg.values()); tots=sum(v[1] for v in agg.values())
print('total inst %.3g samples %d'%(tot,tots))
byf={}
for (f,l),v in agg.items():
    b=byf.setdefault(f,[0,0,0]); b[0]+=v[0]; b[1]+=v[1]; b[2]+=v[4]
for f,b in byf.items(): print('  file',f,'inst %.1f%% smp %.1f%% thr %.1f'%(100*b[0]/tot,100*b[1]/tots,b[2]/max(b[0],1)))
for (f,l),v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:topn]:
    top=sorted(v[3].items(), key=lambda kv:-kv[1])[:3]
    print(f,l,'inst %.2f%% smp %.2f%% thr %.0f'%(100*v[0]/tot,100*v[1]/tots,v[4]/max(v[0],1)), ' '.join('%s=%d'%(k[6:],x) for k,x in top), '|', v[2].strip()[:80])
