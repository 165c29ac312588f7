import csv, sys, collections
fn=sys.argv[1]; which=int(sys.argv[2])
rows=csv.reader(open(fn))
k=-1; cur=[]; sect=[]
for r in rows:
    if r and r[0]=="Kernel Name":
        k+=1; continue
    if r and r[0]=="Address": hdr=r; continue
    if k==which: cur.append(r)
ie=hdr.index("Instructions Executed"); src=hdr.index("Source")
st_cols=[i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot=0; hist=collections.Counter(); stall=collections.Counter(); sthist=collections.defaultdict(collections.Counter)
n_hot=0
mx=max(int(r[ie]) for r in cur)
for r in cur:
    c=int(r[ie]); tot+=c
    op=r[src].split()
    if op and op[0].startswith('@'): op=op[1:]
    o=op[0].split('.')[0] if op else '?'
    hist[o]+=c
    if c>0.2*mx: n_hot+=1
    for i in st_cols:
        v=int(r[i] or 0)
        stall[hdr[i]]+=v; sthist[o][hdr[i]]+=v
print("lines",len(cur),"total warp inst",tot,"max per line",mx,"hot lines(>0.2max)",n_hot)
for o,c in hist.most_common(40): print(f"{o:10s} {c/tot*100:6.2f}%  {c}")
print({k:v for k,v in stall.most_common(12)})
tots=sum(stall.values())
print("stall samples by opcode:")
agg={o:sum(d.values()) for o,d in sthist.items()}
for o,v in sorted(agg.items(), key=lambda x:-x[1])[:15]:
    print(f"{o:10s} {v/tots*100:5.1f}% ", dict(sthist[o].most_common(3)))
