#!/usr/bin/env python
"""scripts/ncu_stall_buckets.py SOURCE_PAGE.csv [WARP_STEPS [BUCKET]] -- warp-stall samples of an `ncu --page source --csv` export summed
over buckets of SASS instructions (share of warp time, executed instructions per warp and step, top stall reasons, opcodes seen)."""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
S='# Samples'
tot=sum(int(r[ix[S]] or 0) for r in data)
nwarp=int(sys.argv[2]) if len(sys.argv)>2 else 114688
print('total',tot,len(data))
agg={s:sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
for k,v in sorted(agg.items(), key=lambda x:-x[1])[:10]: print(f'{k:28s} {v:8d} {100*v/tot:5.1f}%')
B=int(sys.argv[3]) if len(sys.argv)>3 else 100
cum=0
for b in range(0,len(data),B):
    seg=data[b:b+B]
    s=sum(int(r[ix[S]] or 0) for r in seg)
    ex=sum(int(r[ix['Instructions Executed']] or 0) for r in seg)
    if s==0 and ex==0: continue
    cum+=s
    st=sorted(((sum(int(r[ix[k]] or 0) for r in seg),k[6:]) for k in stalls), reverse=True)[:3]
    marks=set()
    for r in seg:
        t=r[ix['Source']]
        for m in ('BAR.SYNC','UCGABAR','MUFU','SHFL','FFMA','LDS','STS','LDG','STG','CALL','RET','SYNCS','LDL','STL','UBLKCP','ST.E','LD.E','STAS','MAPA','NANOSLEEP','ATOMS'):
            if m in t: marks.add(m)
    print(f'{b:5d} {s:6d} {100*s/tot:5.1f}% cum {100*cum/tot:5.1f}% ex/ws {ex/nwarp:6.1f} {st} {sorted(marks)}')
