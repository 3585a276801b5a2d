"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line:
stall samples, executed instructions and average active threads.  usage: ncu_lines.py dump.csv [top_n]"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; cur_file=None
agg=[]
for r in rows:
    if r and r[0]=='File Path': cur_file=r[1]; continue
    if r and r[0]=='Line No': hdr=r; continue
    if hdr and r and r[0].isdigit():
        line=int(r[0]); src=r[1]
        try: agg.append((cur_file.split('/')[-1],line,src.strip()[:90],int(r[hdr.index('# Samples')]),int(r[hdr.index('Instructions Executed')]),int(r[hdr.index('Thread Instructions Executed')])))
        except: pass
tot_s=sum(a[3] for a in agg); tot_i=sum(a[4] for a in agg)
print("total samples",tot_s,"total inst",tot_i)
N=int(sys.argv[2]) if len(sys.argv)>2 else 40
for a in sorted(agg,key=lambda x:-x[3])[:N]:
    print(f"{a[0]}:{a[1]:4d} samp {100*a[3]/tot_s:5.1f}% inst {100*a[4]/tot_i:5.1f}% thr/inst {a[5]/max(a[4],1):5.1f} | {a[2]}")
