import csv, collections, sys
def summarize(path, top=22):
    rows=list(csv.reader(open(path)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
    h=rows[hdr]; data=rows[hdr+1:]
    ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
    agg=collections.OrderedDict()
    for r in data:
        if len(r)<=vi: continue
        v=float(r[vi].replace(",",""))
        if r[ui]=="ns": v/=1e3
        elif r[ui]=="ms": v*=1e3
        k=r[ki][:100]
        a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
    tot=sum(a[1] for a in agg.values())
    print(f"# {path}: total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
    for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:top]:
        print(f"{a[1]:10.1f} us {a[0]:4d}x {100*a[1]/tot:5.1f}%  {k}")
if __name__=="__main__":
    for p in sys.argv[1:]: summarize(p)
