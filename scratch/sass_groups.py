import csv, sys
from collections import Counter
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=[r for r in rows[2:] if len(r)>=len(hdr)-2 and r[0].startswith('0x')]
iS=hdr.index("Source"); iE=hdr.index("Instructions Executed"); iSt=hdr.index("# Samples")
tot=sum(int(r[iE]) for r in data); print("total warp instr", tot, "total samples", sum(int(r[iSt]) for r in data), "n sass", len(data))
groups=[]; cur=None
for k,r in enumerate(data):
    e=int(r[iE]); s=int(r[iSt]); op=r[iS].split()
    op=[o for o in op if not o.startswith('@')]
    op=op[0] if op else ''
    if cur and cur['e']==e: cur['n']+=1; cur['samples']+=s; cur['ops'].append(op)
    else:
        cur={'start':k,'e':e,'n':1,'samples':s,'ops':[op]}; groups.append(cur)
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.01
for g in groups:
    if g['e']*g['n'] > tot*thr:
        c=Counter(o.split('.')[0] for o in g['ops'])
        print(f"idx {g['start']:5d} n={g['n']:4d} exec/inst={g['e']:9d} share={g['e']*g['n']/tot*100:5.1f}% samples={g['samples']:6d}  {dict(c.most_common(10))}")
