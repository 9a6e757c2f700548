#!/usr/bin/env python3
"""ncu source page (ncu -i rep --page source --csv --print-source sass,cuda > x.csv) -> opcode mix, stall mix and the hottest source lines.
usage: ncu_source_hotspots.py x.csv [top_n]"""
import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
cur=None; curline=None
byline=collections.OrderedDict(); ops=collections.Counter(); opsamp=collections.Counter(); stall=collections.Counter()
tot_i=tot_s=0
hdr=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; ie=hdr.index("Instructions Executed"); isamp=hdr.index("# Samples"); st_cols=[(i,n) for i,n in enumerate(hdr) if n.startswith('stall_') and 'Not Issued' not in n]; continue
    if r[0]!='' :  # cuda source line row
        curline=(cur,int(r[0]),r[1].strip()[:90]); byline.setdefault(curline,[0,0]); continue
    try: n=int(r[ie]); s=int(r[isamp])
    except: continue
    sass=r[3].split()
    op=sass[1] if sass and sass[0].startswith('@') else (sass[0] if sass else '?')
    op=op.split('.')[0]
    ops[op]+=n; opsamp[op]+=s; tot_i+=n; tot_s+=s
    byline[curline][0]+=n; byline[curline][1]+=s
    for i,nm in st_cols:
        try: stall[nm]+=int(r[i])
        except: pass
print('total warp-inst',tot_i,'samples',tot_s)
print('stalls:',[(k,round(100*v/tot_s,1)) for k,v in stall.most_common(8)])
for op,n in ops.most_common(18): print('%-8s inst %6.2f%%  samples %6.2f%%'%(op,100*n/tot_i,100*opsamp[op]/tot_s))
print('--- top lines by samples')
for k,(n,s) in sorted(byline.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print('%5.2f%% smp %5.2f%% inst  %s:%d  %s'%(100*s/tot_s,100*n/tot_i,k[0],k[1],k[2]))
