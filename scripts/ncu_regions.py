"""Attribute warp-stall samples of one kernel to code regions (split at barriers / hot branches)."""
import csv, subprocess, io, sys
rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name",f"regex:{kern}"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="Address")
hdr=rows[hi]; col={h:i for i,h in enumerate(hdr)}
def num(r,k):
    try: return float(r[col[k]])
    except Exception: return 0.0
data=[];seen=set()
for r in rows[hi+1:]:
    if len(r)!=len(hdr) or r[col['Address']] in seen or r[col['Address']]=='Address': continue
    seen.add(r[col['Address']]); data.append(r)
tot=sum(num(r,'# Samples') for r in data)
stalls=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
acc=0; start=0; sacc={s:0 for s in stalls}; iex=0
def flush(i,label):
    global acc,start,sacc,iex
    if acc/tot>0.004:
        top=sorted(sacc.items(),key=lambda x:-x[1])[:3]
        print(f"insts {start:5d}-{i:5d} {acc/tot*100:5.1f}% exec/inst {iex/max(1,i-start+1)/1e6:7.2f}M  {label:42s} "+" ".join(f"{k[6:]}={v/acc*100:.0f}%" for k,v in top))
    acc=0; start=i+1; sacc={s:0 for s in stalls}; iex=0
for i,r in enumerate(data):
    acc+=num(r,'# Samples'); iex+=num(r,'Instructions Executed')
    for s in stalls: sacc[s]+=num(r,s)
    src=r[col['Source']]
    if 'BAR.' in src or ('BRA' in src) or 'WARPSYNC' in src: flush(i,'up to '+src.strip()[:36])
flush(len(data)-1,'end')
