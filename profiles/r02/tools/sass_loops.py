"""Static SASS summary of a kernel: every innermost loop of >= 100 instructions with its opcode histogram.

    python profiles/r02/tools/sass_loops.py curiousrl_b200/libcuriousrl_b200.so [mangled-name substring]

Used on the CPU (cuobjdump only) to see what the time-step loops of the solve kernel compile to before spending GPU time.
"""
import re, sys, collections, subprocess
so = sys.argv[1]; fn_pat = sys.argv[2] if len(sys.argv)>2 else 'solve_kernelINS_9ThreeLinkEdLi4ELi2ELb1E'
sass = subprocess.run(['cuobjdump','-sass',so],capture_output=True,text=True).stdout
lines = sass.split('\n')
start=end=None
for i,l in enumerate(lines):
    if 'Function :' in l:
        if start is not None and end is None: end=i
        if fn_pat in l and start is None and ('coop_solve' in fn_pat or 'coop_solve' not in l): start=i; end=None
if end is None: end=len(lines)
pat = re.compile(r'^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\*')
ins=[]
for l in lines[start:end]:
    m=pat.match(l)
    if m: ins.append((int(m.group(1),16), m.group(2).strip()))
def opcode(t):
    t = re.sub(r'^@!?U?P\d+\s+', '', t)
    return t.split()[0].split('.')[0] if t else '?'
a2i={a:i for i,(a,_) in enumerate(ins)}
loops=[]
for i,(a,t) in enumerate(ins):
    if re.search(r'\bBRA\b',t):
        m=re.search(r'0x([0-9a-f]+)',t)
        if m:
            tg=int(m.group(1),16)
            if tg<=a and tg in a2i: loops.append((a2i[tg],i))
loops.sort()
print('total instr',len(ins))
inner=[(s,e) for (s,e) in loops if e-s>=100 and not any((s2>=s and e2<=e and (s2,e2)!=(s,e) and e2-s2>=100) for (s2,e2) in loops)]
for s,e in inner:
    h=collections.Counter(opcode(t) for _,t in ins[s:e+1]); n=e-s+1
    fp=sum(v for k,v in h.items() if k in('DFMA','DADD','DMUL','DSETP'))
    tag='?'
    if h['SHFL']>=60: tag='colgroup'
    elif h['LDS']>=40 and h['STS']>=10: tag='elem-per-lane'
    elif h['LDG']+h['LD']>=25 and h['DFMA']>100: tag='coop_rollout/recurrence'
    elif n>1800 and h['DFMA']>300: tag='multi-alpha'
    elif n>1100 and h['DFMA']>300: tag='lane-backward'
    elif h['LDS']>=25 and h['DFMA']>90: tag='lane-rollout1'
    elif h['DFMA']>80 and h['STG']>=10: tag='open-rollout'
    print('0x%05x n=%4d fp64=%3d %-24s %s'%(ins[s][0],n,fp,tag,' '.join('%s:%d'%kv for kv in h.most_common(14))))
