import csv, sys
from collections import Counter
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=[r for r in rows[2:] if len(r)>=len(hdr)-2 and r[0].startswith('0x')]
n=len(data) if len(sys.argv)>4 else len(data)//2; data=data[:n]
iS=hdr.index("Source"); iE=hdr.index("Instructions Executed"); iSt=hdr.index("# Samples")
tot=sum(int(r[iE]) for r in data); ts=sum(int(r[iSt]) for r in data)
ncta=int(sys.argv[2]); print("warp-instr per CTA", tot/ncta, "n sass", n)
B=int(sys.argv[3]) if len(sys.argv)>3 else 100
for b in range(0,n,B):
    blk=data[b:b+B]
    e=sum(int(r[iE]) for r in blk); s=sum(int(r[iSt]) for r in blk)
    c=Counter()
    for r in blk:
        op=[o for o in r[iS].split() if not o.startswith('@')]
        if op: c[op[0].split('.')[0]]+=int(r[iE])
    top=", ".join(f"{k}:{v/ncta:.0f}" for k,v in c.most_common(7))
    print(f"[{b:4d}] instr/CTA={e/ncta:8.0f} ({e/tot*100:4.1f}%) samples={s/ts*100:4.1f}%  {top}")
# total by opcode
c=Counter()
for r in data:
    op=[o for o in r[iS].split() if not o.startswith('@')]
    if op: c[op[0].split('.')[0]]+=int(r[iE])
print("TOTAL by opcode per CTA:", ", ".join(f"{k}:{v/ncta:.0f}" for k,v in c.most_common(25)))
