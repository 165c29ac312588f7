import csv, sys, re, collections
sass, a, b, prof, which = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
lines=open(sass).read().split('\n')[a-1:b]
ins=[]; cur=None
for l in lines:
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        # keep the outermost?? use innermost (this) plus inlined-at chain
        cur=(m.group(1).split('/')[-1], int(m.group(2)), m.group(3))
        continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: ins.append((cur, m.group(2)))
rows=csv.reader(open(prof)); k=-1; curk=[]
for r in rows:
    if r and r[0]=="Kernel Name": k+=1; continue
    if r and r[0]=="Address": hdr=r; continue
    if k==which: curk.append(r)
print(len(ins), len(curk))
ie=hdr.index("Instructions Executed"); ss=hdr.index("Warp Stall Sampling (All Samples)")
by=collections.defaultdict(lambda:[0,0])
n=min(len(ins),len(curk))
for (loc,op),r in zip(ins[:n],curk[:n]):
    key=(loc[0],loc[1]) if loc else None
    by[key][0]+=int(r[ie]); by[key][1]+=int(r[ss] or 0)
ti=sum(v[0] for v in by.values()); ts=sum(v[1] for v in by.values())
print("by executed instructions")
for k,v in sorted(by.items(), key=lambda x:-x[1][0])[:45]:
    print(k, f"{v[0]/ti*100:5.2f}% inst  {v[1]/ts*100:5.2f}% stall")
