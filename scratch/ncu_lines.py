"""Per-source-line attribution of an ncu report: python scratch/ncu_lines.py report.ncu-rep [kernel-index] [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
kern = -1; agg = {}; fileName = None; hdr = None; names = []
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name": kern += 1; names.append(r[1]); continue
    if r[0] == "File Path": fileName = r[1].split('/')[-1]; continue
    if r[0] == "Function Name":
        if kern < 0 or len(names) == 0 or True: pass
        continue
    if r[0] == "Line No": hdr = r; continue
    if r[0].isdigit() and hdr:
        iI = hdr.index('Instructions Executed'); iSm = hdr.index('# Samples'); iW = hdr.index('L1 Wavefronts Shared')
        key = (kern, fileName, int(r[0]))
        a = agg.setdefault(key, [r[1], 0, 0, 0])
        a[1] += num(r[iI]); a[2] += num(r[iSm]); a[3] += num(r[iW])
ks = sorted(set(k[0] for k in agg))
print("kernels:", names if names else ks)
sel = {k: v for k, v in agg.items() if k[0] == (ks[which] if ks else -1)}
tot = sum(v[1] for v in sel.values()) or 1
print("total warp instr", tot)
for k, v in sorted(sel.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[1]:20s} {k[2]:4d} {v[1]/tot*100:5.1f}% smp {v[2]:6d} wf {v[3]:10d}  {v[0][:100]}")
